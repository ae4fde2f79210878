#!/usr/bin/env python
"""bench.py — one denoise-step forward of the Wan2.2-5B FrameINO transformer on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--frames 121]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

A "step" is ONE forward of boundary #1 (reference pipeline: 2 per scheduler step with CFG). Workload = BASELINE.json
configs[1]: 704x1280x121 canvas -> 27280 video + 880 ID tokens = 28160 tokens, D 3072, 24 heads x 128, FFN 14336,
30 layers, 512 text tokens, per-token timesteps; synthetic seeded latents, random-init weights.
N > 1: Ulysses sequence parallel over the token axis, strong scaling (the same step is sharded).

The JSON line carries: value (ms, device-resident inputs, max over ranks), e2e (same step through the public
forward() with pinned-host inputs copied in and the output copied out inside the timed region), roofline of the
dominant kernel (tcgen05 flash attention: algorithmic 4*N^2*D FLOPs / mean CUDA-event duration of the launches inside
the timed steps, against the measured sustained bf16 peak), parity (N = 1: the native model cut to block 0 on the bench
inputs vs the CPU oracle on the same weights, every tap; N > 1: the sharded forward vs the same forward un-sharded on
rank 0), cpu_baseline (that same oracle run, timed, on this box's host cores — a bounded sample, extrapolated over
the 30 identical layers and labelled so), clocks sampled during the timed region, gpu_launches (kernels of
libframeino_b200.so launched in the timed region), and secondary (after the headline timing: BASELINE config 3,
CogVideoX-5B B = 2; at N = 8 also config 5, the 39936-token canvas).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "wan22_5b_frameino_denoise_step_ms"
UNIT = "ms"


def workload(frames: int, height: int = 704, width: int = 1280):
    """Latent geometry of a height x width canvas with `frames` pixel frames + 1 ID frame (SURVEY.md §8d config 2;
    832 x 1536 is config 5, the expanded canvas)."""
    lat_f = (frames - 1) // 4 + 1
    h, w = height // 16, width // 16
    tokens = (lat_f + 1) * (h // 2) * (w // 2)
    return lat_f, h, w, tokens


def flops_per_forward(tokens: int, d=3072, ffn=14336, text=512, layers=30, c_in=96, c_out=48) -> float:
    per_layer = (3 * 2 * tokens * d * d + 4 * tokens * tokens * d + 2 * tokens * d * d + 4 * tokens * d * d +
                 4 * text * d * d + 4 * tokens * text * d + 4 * tokens * d * ffn)
    return layers * per_layer + 2 * tokens * (c_in * 4) * d + 2 * tokens * d * (c_out * 4) + 2 * text * (4096 + d) * d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg, parity leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_dtype():
    """bf16 (the config's dtype, the reference's cast points live) when this host multiplies bf16 at speed (AMX /
    AVX512-BF16), else fp32. Returns (torch dtype, probe TFLOP/s of both)."""
    import torch

    a32 = torch.randn(1024, 2048)
    b32 = torch.randn(2048, 2048)
    res = {}
    for dt in (torch.float32, torch.bfloat16):
        a, b = a32.to(dt), b32.to(dt)
        torch.nn.functional.linear(a, b)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.linear(a, b)
        res[dt] = 3 * 2 * 1024 * 2048 * 2048 / (time.perf_counter() - t0) / 1e12
    dt = torch.bfloat16 if res[torch.bfloat16] >= res[torch.float32] else torch.float32
    return dt, {"f32_tflops": round(res[torch.float32], 3), "bf16_tflops": round(res[torch.bfloat16], 3)}


class OracleForward:
    """The CPU oracle (oracle/wan_oracle.py) on the bench workload with ONE of the 30 identical full-width blocks:
    prologue (RoPE tables, patch embedding, the reference's per-token time MLP, text MLP), block 0 (D 3072, FFN 14336,
    24 x 128, 512 text tokens), epilogue (output modulation, proj_out, un-patchify) over `frames_s` + 1 latent frames
    (all 31 + 1 = the full 28160 tokens when it fits the time budget). The parts are timed separately; a full 30-layer
    forward is prologue + 30 x block + epilogue (the layers do identical work), and when a token sub-sample is used the
    token-linear parts scale by N/N_s and the self-attention core by (N/N_s)^2 — an EXTRAPOLATION either way, and the
    JSON says so."""

    def __init__(self, lat_f, h, w, dtype, sd=None, inputs=None):
        import torch

        from frameino_b200 import synth
        from oracle import wan_oracle

        self.torch, self.wo, self.synth = torch, wan_oracle, synth
        cfg = dict(synth.WAN22_5B)
        cfg["num_layers"] = 1
        self.cfg_dict = cfg
        self.cfg = wan_oracle.WanConfig(**cfg)
        self.dtype = dtype
        self.lat_f, self.h, self.w = lat_f, h, w
        self.sd = sd if sd is not None else synth.make_wan_state_dict(cfg, seed=0, dtype=dtype)
        self.inputs = inputs
        self.per_frame = (h // 2) * (w // 2)
        self.tokens = (lat_f + 1) * self.per_frame

    def make_inputs(self, frames_s):
        if self.inputs is not None and frames_s == self.lat_f:
            return self.inputs
        return self.synth.make_wan_inputs(self.cfg_dict, frames_s, self.h, self.w, n_id=1, text_len=512,
                                          text_true_len=120, dtype=self.dtype)

    def run(self, frames_s, taps=None):
        """One oracle forward with a single block over (frames_s + 1) latent frames. Returns (sample, timing dict)."""
        torch, wo = self.torch, self.wo
        hidden, ts, text = self.make_inputs(frames_s)
        t_attn, t_block = [0.0], [0.0]
        real_sdpa, real_block = wo.sdpa, wo.wan_block

        def timed_sdpa(q, k, v):
            t0 = time.perf_counter()
            o = real_sdpa(q, k, v)
            if q.shape[2] == k.shape[2]:  # self-attention core only
                t_attn[0] += time.perf_counter() - t0
            return o

        def timed_block(*a, **k):
            t0 = time.perf_counter()
            o = real_block(*a, **k)
            t_block[0] += time.perf_counter() - t0
            return o

        wo.sdpa, wo.wan_block = timed_sdpa, timed_block
        try:
            t0 = time.perf_counter()
            with torch.no_grad():
                out = wo.wan_forward(self.sd, self.cfg, hidden, ts, text, taps=taps, num_layers=1)
            total = time.perf_counter() - t0
        finally:
            wo.sdpa, wo.wan_block = real_sdpa, real_block
        n_s = (frames_s + 1) * self.per_frame
        r = self.tokens / n_s
        outside = total - t_block[0]
        lin = t_block[0] - t_attn[0]
        est = (outside * r + 30.0 * (lin * r + t_attn[0] * r * r)) * 1e3
        return out, {"tokens": n_s, "total_s": total, "block_s": t_block[0], "attn_core_s": t_attn[0],
                     "outside_block_s": outside, "forward_ms_extrapolated": est}

    def describe(self, frames_s, n_runs) -> str:
        n_s = (frames_s + 1) * self.per_frame
        dt = "bf16 weights/activations with the reference's fp32 islands" if self.dtype == self.torch.bfloat16 else "fp32"
        how = ("all tokens: prologue + 30 x block + epilogue (extrapolated over the 30 identical layers only)"
               if n_s == self.tokens else
               f"token sub-sample: token-linear time x{self.tokens / n_s:.2f}, self-attention core x{(self.tokens / n_s) ** 2:.2f}, "
               "block x30 (extrapolated)")
        return (f"CPU oracle (oracle/wan_oracle.py, {dt}): prologue + 1 of 30 full-width blocks + epilogue over {n_s} of "
                f"{self.tokens} tokens, {n_runs} run(s); {how}")


def run_reference(args, lat_f, h, w, tokens):
    """--impl reference: the reference's CPU implementation of the path. The reference itself cannot be installed
    here (needs diffusers, absent from the image and the wheelhouse), so this times the oracle port on all host cores.
    Every step is a bounded sample (one of the 30 blocks + prologue/epilogue), the first warm-up step always over ALL
    tokens; the remaining steps use all tokens too when (steps + warmup) of them fit ~4 minutes, else a frame sub-sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use the whole host
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    dtype, probe = cpu_oracle_dtype()
    orc = OracleForward(lat_f, h, w, dtype)
    _, full = orc.run(lat_f)  # calibration: one block over all tokens, measured not extrapolated
    n_total = args.steps + max(args.warmup, 1)
    budget_s = float(os.environ.get("FINO_REF_BUDGET_S", "240"))
    frames_s = lat_f
    if full["total_s"] * (n_total - 1) > budget_s:
        per = full["total_s"]
        for cand in (15, 7, 3):  # (cand + 1) / 32 of the tokens
            frames_s = cand
            r = (cand + 1) / (lat_f + 1)
            per = full["outside_block_s"] * r + (full["block_s"] - full["attn_core_s"]) * r + full["attn_core_s"] * r * r
            if per * (n_total - 1) <= budget_s:
                break
    for _ in range(max(args.warmup, 1) - 1):
        orc.run(frames_s)
    runs = [orc.run(frames_s)[1] for _ in range(args.steps)]
    v = sum(r["forward_ms_extrapolated"] for r in runs) / len(runs)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if dtype == torch.bfloat16 else "f32", "data": "synthetic",
        "config": {"workload": f"Wan2.2-TI2V-5B FrameINO one denoise-step forward, {args.height}x{args.width}x{args.frames} + 1 ID frame, "
                               f"{tokens} tokens, B=1 (CPU oracle port; value EXTRAPOLATED from one of the 30 blocks, "
                               "see cpu_baseline.sample)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
                         "sample": orc.describe(frames_s, len(runs)),
                         "measured_block_all_tokens_s": round(full["block_s"], 3),
                         "measured_prologue_epilogue_all_tokens_s": round(full["outside_block_s"], 3),
                         "forward_ms_from_all_token_block": round(full["forward_ms_extrapolated"], 1),
                         "host_matmul_probe": probe},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def native_block_parity(model, d_in, lat_f, h, w, dev):
    """Parity AT the benchmarked configuration (N = 1): the native model, cut to its first block, runs the bench inputs
    (all tokens, full width) and every tapped tensor is compared with the CPU oracle evaluated on the SAME weights
    (copied off the device) — north-star bar: per-tensor max|a-b| / max|b| <= 2e-2, cosine >= 0.999. The oracle run is
    also the cpu_baseline leg (its parts are timed)."""
    import torch
    from torch import nn

    keep = {k: v.detach().cpu() for k, v in model.state_dict().items()
            if not k.startswith("blocks.") or k.startswith("blocks.0.")}
    host_in = tuple(t.cpu() for t in d_in)
    dtype, probe = cpu_oracle_dtype()
    if dtype != torch.bfloat16:  # no fast bf16 on this host: fp32 oracle on the same (bf16-valued) weights
        keep = {k: v.float() for k, v in keep.items()}
        host_in = tuple(t.float() for t in host_in)
    orc = OracleForward(lat_f, h, w, dtype, sd=keep, inputs=host_in)
    ref_taps = {}
    ref, timing = orc.run(lat_f, taps=ref_taps)

    blocks = model.blocks
    taps = {}
    try:
        model.blocks = nn.ModuleList([blocks[0]])
        model._sst_cache = None
        model.__dict__["_fino_taps"] = taps
        out = model(hidden_states=d_in[0], timestep=d_in[1], encoder_hidden_states=d_in[2], return_dict=False)[0]
        torch.cuda.synchronize()
    finally:
        model.blocks = blocks
        model._sst_cache = None
        model.__dict__.pop("_fino_taps", None)

    def rel(a, b):
        a, b = a.float().cpu(), b.float()
        return float((a - b).abs().max() / b.abs().max())

    names = ["patch_embed", "text", "blocks.0.norm1", "blocks.0.after_attn1", "blocks.0.after_attn2", "blocks.0.out"]
    per_tap = {n: rel(taps[n], ref_taps[n]) for n in names}
    per_tap["sample"] = rel(out, ref)
    cos = float(torch.nn.functional.cosine_similarity(out.float().cpu().flatten(), ref.float().flatten(), dim=0))
    parity = {"rel_err": max(per_tap.values()), "cosine": cos, "per_tap": {k: round(v, 5) for k, v in per_tap.items()},
              "what": f"native model cut to block 0 (full width: D 3072, FFN 14336, 24x128, per-token modulation) on the "
                      f"bench inputs, all {orc.tokens} tokens, vs the CPU oracle on the same weights "
                      f"({'bf16 with fp32 islands' if dtype == torch.bfloat16 else 'fp32'}); rel_err = max over the taps of "
                      "max|a-b| / max|b|, cosine of the output sample; bar 2e-2 / 0.999",
              "pass": bool(max(per_tap.values()) <= 2e-2 and cos >= 0.999)}
    cpu_leg = {"value": timing["forward_ms_extrapolated"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "extrapolated": True, "sample": orc.describe(lat_f, 1),
               "measured_block_all_tokens_s": round(timing["block_s"], 3),
               "measured_prologue_epilogue_all_tokens_s": round(timing["outside_block_s"], 3),
               "host_matmul_probe": probe}
    return parity, cpu_leg


def sharded_vs_unsharded(model, fn, y_sharded, rank):
    """Rank 0 switches sequence parallelism off, re-runs ``fn`` (no collectives on that path) and returns
    {"rel_err", "cosine"} of the sharded output against it; other ranks return None (the caller barriers)."""
    import torch

    from frameino_b200.ulysses import _self_attention_modules

    if rank != 0:
        return None
    sp_saved = model.sequence_parallel
    model.sequence_parallel = None
    for a in _self_attention_modules(model):
        a.__dict__.pop("_fino_sp", None)
    try:
        y_single = fn()
        torch.cuda.synchronize()
    finally:
        model.sequence_parallel = sp_saved
        for a in _self_attention_modules(model):
            a.__dict__["_fino_sp"] = sp_saved
    a32, b32 = y_sharded.float().flatten(), y_single.float().flatten()
    return {"rel_err": float((a32 - b32).abs().max() / b32.abs().max()),
            "cosine": float(torch.nn.functional.cosine_similarity(a32, b32, dim=0))}


def _time_steps(fn, steps, warmup, barrier):
    import torch

    for _ in range(warmup):
        out = fn()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        out = fn()
    e.record()
    barrier()
    return s.elapsed_time(e) / steps, out


def run_secondary(args, world, rank, dev, wan_model, barrier):
    """Secondary workloads after the headline timing, so that they get driver-run numbers (not the headline metric):
    BASELINE config 3 — CogVideoX-5B-I2V FrameINO one denoise step, 480x720x49 + 1 ID frame, S = 19126, B = 2 (the
    pipeline's batched CFG): on 1 GPU, and sequence-parallel when N > 1; at N = 8 also config 5 — the Wan forward on the
    expanded 832x1536x121 canvas (39936 tokens)."""
    import torch
    import torch.distributed as dist

    from frameino_b200 import synth

    out = {}
    steps, warmup = 3, 2
    if world == 8:
        lat_f, h, w, tokens = workload(121, 832, 1536)
        hidden, ts, text = synth.make_wan_inputs(synth.WAN22_5B, lat_f, h, w, n_id=1, text_len=512, text_true_len=120,
                                                 dtype=torch.bfloat16)
        d5 = [t.to(dev) for t in (hidden, ts, text)]
        fn5 = lambda: wan_model(hidden_states=d5[0], timestep=d5[1], encoder_hidden_states=d5[2], return_dict=False)[0]  # noqa: E731
        ms, y = _time_steps(fn5, steps, warmup, barrier)
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cmp5 = sharded_vs_unsharded(wan_model, fn5, y, rank)
        barrier()
        out["config5_wan_40k_tokens"] = {
            "metric": METRIC, "value": float(t.item()), "unit": "ms", "n_gpus": world, "steps": steps, "warmup": warmup,
            "finite": bool(torch.isfinite(y.float()).all()), "parity_sharded_vs_unsharded": cmp5,
            "config": {"workload": f"Wan2.2-TI2V-5B FrameINO one denoise-step forward, 832x1536x121 + 1 ID frame, {tokens} "
                                   f"tokens, B=1, ulysses x{world} ({args.sp_mode})"}}
        del d5, y
    # config 3 (sequence-parallel when N > 1; --no-cog-sp skips it there)
    if world > 1 and args.no_cog_sp:
        return out
    try:
        cfg = synth.COG_5B_I2V
        cog = synth.build_cog_on_device(cfg, seed=0, device=dev)
        if world > 1:
            from frameino_b200.ulysses import enable_sequence_parallel

            enable_sequence_parallel(cog, mode=args.cog_sp_mode)
        lat_f, h, w, batch = 13, 60, 90, 2
        hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=batch, dtype=torch.bfloat16)
        cos, sin = synth.cog_rope_tables(64, h // 2, w // 2, lat_f, 1, device=dev)
        dc = [hidden.to(dev), text.to(dev), ts.to(dev)]
        seq = 226 + (lat_f + 1) * (h // 2) * (w // 2)
        ms, y = _time_steps(lambda: cog(hidden_states=dc[0], encoder_hidden_states=dc[1], timestep=dc[2],
                                        image_rotary_emb=(cos, sin), return_dict=False)[0], steps, warmup, barrier)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        d, f, layers = 3072, 12288, 42
        flops = batch * layers * (3 * 2 * seq * d * d + 4 * seq * seq * d + 2 * seq * d * d + 4 * seq * d * f)
        out["config3_cogvideox_b2"] = {
            "metric": "cogvideox_5b_i2v_frameino_denoise_step_ms", "value": ms, "unit": "ms", "n_gpus": world,
            "steps": steps, "warmup": warmup, "finite": bool(torch.isfinite(y.float()).all()),
            "model_tflops": flops / (ms * 1e-3) / 1e12,
            "config": {"workload": f"CogVideoX-5B-I2V FrameINO one denoise step, 480x720x49 + 1 ID frame, S={seq}, B={batch}"
                                   + ("" if world == 1 else f", ulysses x{world} ({args.cog_sp_mode})")}}
        if cog.sequence_parallel is not None:
            cog.sequence_parallel.close()
        del cog, dc, y
    except Exception as ex:  # the secondary line must never take the headline down
        out["config3_cogvideox_b2"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    torch.cuda.empty_cache()
    if world == 1:
        try:
            out["wan_vae_704x1280x121"] = run_vae_secondary(dev)
        except Exception as ex:
            out["wan_vae_704x1280x121"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()
    else:
        try:
            out["wan_vae_704x1280x121_row_parallel"] = run_vae_rows_secondary(dev, world, barrier)
        except Exception as ex:
            out["wan_vae_704x1280x121_row_parallel"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()
    return out


def run_vae_rows_secondary(dev, world, barrier):
    """N > 1: the Wan2.2 VAE split by frame rows over the ranks (frameino_b200/vae.py RowParallel; halo rows through peer
    mailboxes) at config 2's canvas: decode and encode, sharded vs the un-sharded run of the same model on the same
    inputs (expected bit-identical), device times as the max over ranks. Same sequence as tools/vae_sp_check.py --full."""
    import torch
    import torch.distributed as dist

    from frameino_b200 import synth

    vae = synth.build_vae_on_device(synth.WAN22_VAE, seed=0, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    z = torch.randn(1, 48, 31, 44, 80, generator=g, device=dev)

    def timed(fn):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return y, float(t.item())

    def max_diff(a, b):
        d = (a - b).abs().max().reshape(1).float()
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        return float(d.item())

    vae.decode(z[:, :, :2], return_dict=False)
    ref, dec_1 = timed(lambda: vae.decode(z, return_dict=False)[0])
    vae.enable_row_parallel()
    vae.decode(z[:, :, :2], return_dict=False)
    got, dec_n = timed(lambda: vae.decode(z, return_dict=False)[0])
    dec_diff = max_diff(got, ref)
    del got, ref
    x = torch.randn(1, 3, 121, 704, 1280, generator=g, device=dev).clamp_(-1, 1)
    vae.encode(x[:, :, :5])
    enc_rp, enc_n = timed(lambda: vae.encode(x).latent_dist.parameters)
    vae.disable_row_parallel()
    vae.encode(x[:, :, :5])
    enc, enc_1 = timed(lambda: vae.encode(x).latent_dist.parameters)
    enc_diff = max_diff(enc_rp, enc)
    return {"config": {"workload": f"Wan2.2-TI2V-5B VAE (random init), 704x1280x121 <-> latent 31x44x80x48, B=1, frame rows "
                                   f"split over {world} ranks (halo rows through peer memory), inputs resident in HBM"},
            "unit": "ms", "n_gpus": world, "decode_ms": dec_n, "decode_unsharded_ms": dec_1, "encode_ms": enc_n,
            "encode_unsharded_ms": enc_1,
            "parity_sharded_vs_unsharded": {"decode_max_abs_diff": dec_diff, "encode_max_abs_diff": enc_diff,
                                            "pass": bool(dec_diff == 0.0 and enc_diff == 0.0)}}


def run_vae_secondary(dev):
    """SURVEY §8f row 3 beside the loop: the Wan2.2 VAE (random-init TI2V-5B VAE architecture) at config 2's canvas —
    one decode of the [1, 48, 31, 44, 80] latent and one encode of a 704x1280x121 clip, device-resident, CUDA events;
    plus a parity check of the same code path against the CPU oracle at the tiny config (checker only)."""
    import torch

    from frameino_b200 import synth

    vae = synth.build_vae_on_device(synth.WAN22_VAE, seed=0, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    z = torch.randn(1, 48, 31, 44, 80, generator=g, device=dev)
    x = torch.randn(1, 3, 121, 704, 1280, generator=g, device=dev).clamp_(-1, 1)
    res = {"config": {"workload": "Wan2.2-TI2V-5B VAE (704.7 M parameters, random init), 704x1280x121 <-> latent "
                                  "31x44x80x48, B=1, bf16 activations / fp32 accumulation, inputs resident in HBM"},
           "unit": "ms"}
    for name, fn in (("decode", lambda: vae.decode(z, return_dict=False)[0]),
                     ("encode", lambda: vae.encode(x).latent_dist.mode())):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = fn()
        e.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = s.elapsed_time(e)
        res[name + "_finite"] = bool(torch.isfinite(y.float()).all())
        del y
    del vae, z, x
    torch.cuda.empty_cache()
    # parity of the same kernels / chunking against the oracle, at a size the CPU finishes in a second
    from frameino_b200.vae import AutoencoderKLWan
    from oracle import vae_oracle

    cfg = synth.VAE_TINY
    sd = synth.make_vae_state_dict(cfg, seed=1)
    tiny = AutoencoderKLWan(**cfg)
    tiny.load_state_dict(sd, strict=True)
    tiny = tiny.to(dev).eval().prepare()
    zt, xt = synth.make_vae_inputs(cfg, latent_frames=3, h=4, w=6)
    taps = {}
    vae_oracle.decode(sd, cfg, zt, taps)
    want_enc = vae_oracle.encode(sd, cfg, xt)[:, : cfg["z_dim"]]
    tiny.__dict__["_fino_taps"] = {}
    tiny.decode(zt.to(dev), return_dict=False)
    got_head = tiny.__dict__.pop("_fino_taps")["head"]
    got_enc = tiny.encode(xt.to(dev)).latent_dist.mode()

    def rel(a, b):
        return float((a.float().cpu() - b).abs().max() / b.abs().max())

    head = taps["head"]
    got_head = torch.cat(got_head, dim=0).permute(3, 0, 1, 2)[None]  # per-chunk channels-last frames -> NCTHW
    res["parity"] = {"what": "tiny VAE config (3 latent frames, 64x96): decoder head before the clamp and posterior mean "
                             "vs the CPU oracle, max |a-b| / max |b|", "decode_head_rel_err": rel(got_head, head),
                     "encode_mean_rel_err": rel(got_enc, want_enc), "tolerance": 2e-2}
    res["parity"]["pass"] = bool(res["parity"]["decode_head_rel_err"] <= 2e-2
                                 and res["parity"]["encode_mean_rel_err"] <= 2e-2)
    return res


# ---------------------------------------------------------------------------------------------------------------
def run_native(args, lat_f, h, w, tokens):
    import torch
    import torch.distributed as dist

    from frameino_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device: frameino_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.WAN22_5B
    model = synth.build_wan_on_device(cfg, seed=0, device=dev)
    if world > 1:
        from frameino_b200.ulysses import enable_sequence_parallel

        enable_sequence_parallel(model, mode=args.sp_mode)
    hidden, ts, text = synth.make_wan_inputs(cfg, lat_f, h, w, n_id=1, text_len=512, text_true_len=120,
                                             dtype=torch.bfloat16)
    host = [t.pin_memory() for t in (hidden, ts, text)]
    d_in = [t.to(dev) for t in host]
    out_host = torch.empty(1, cfg["out_channels"], lat_f + 1, h, w, dtype=torch.bfloat16).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = out_host.numel() * out_host.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return model(hidden_states=d_in[0], timestep=d_in[1], encoder_hidden_states=d_in[2], return_dict=False)[0]

    def step_e2e():
        xs = [t.to(dev, non_blocking=True) for t in host]
        y = model(hidden_states=xs[0], timestep=xs[1], encoder_hidden_states=xs[2], return_dict=False)[0]
        out_host.copy_(y, non_blocking=True)
        return y

    # attention launches are bracketed with CUDA events on the launching stream during the timed steps
    attn_events = []
    real_attention = ops.attention

    def timed_attention(q, k, v, heads, scale=None, out=None):
        if q.shape[1] != k.shape[1]:  # cross-attention (512 keys): not the dominant kernel
            return real_attention(q, k, v, heads, scale=scale, out=out)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        o = real_attention(q, k, v, heads, scale=scale, out=out)
        e.record()
        attn_events.append((s, e, q.shape[1], k.shape[1], heads, q.shape[2] // heads))
        return o

    real_attention_scatter = ops.attention_scatter

    def timed_attention_scatter(q, k, v, heads, *rest):  # the same kernel with the peer-scatter epilogue (N > 1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_attention_scatter(q, k, v, heads, *rest)
        e.record()
        attn_events.append((s, e, q.shape[1], k.shape[1], heads, q.shape[2] // heads))

    def timed_region(fn, steps, instrument):
        barrier()
        if instrument:
            ops.attention = timed_attention
            if model.sequence_parallel is not None:
                model.sequence_parallel._attention = timed_attention
                model.sequence_parallel._attention_scatter = timed_attention_scatter
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        if instrument:
            torch.cuda.nvtx.range_push("fino_timed")  # ncu --nvtx --nvtx-include "fino_timed/": the launch list of the steps
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        if instrument:
            torch.cuda.nvtx.range_pop()
        barrier()
        ops.attention = real_attention
        if model.sequence_parallel is not None:
            model.sequence_parallel._attention = real_attention
            model.sequence_parallel._attention_scatter = real_attention_scatter
        ms = s.elapsed_time(e) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        step_device()
    with ClockSampler(local_rank) as clocks:
        ms, launches = timed_region(step_device, args.steps, instrument=True)
    attn_ms = [s.elapsed_time(e) for (s, e, *_r) in attn_events]
    nq, nk, hh, hd = attn_events[0][2:] if attn_events else (tokens, tokens, 24, 128)
    attn_flops = 4.0 * nq * nk * hh * hd
    attn_mean_ms = sum(attn_ms) / max(len(attn_ms), 1)
    step_e2e()
    barrier()
    e2e_ms, _ = timed_region(step_e2e, args.steps, instrument=False)

    # ---- parity of the configuration that was just timed (outside the timed regions) ---------------------------
    y_timed = step_device()
    torch.cuda.synchronize()
    if not bool(torch.isfinite(y_timed.float()).all()):
        raise AssertionError("the timed forward produced non-finite values")
    parity = {}
    if world > 1:
        # rank 0 re-runs the SAME forward un-sharded (no collectives on that path) and compares
        cmp = sharded_vs_unsharded(model, step_device, y_timed, rank)
        if cmp is not None:
            cmp["what"] = (f"full 30-layer forward, {world}-way Ulysses ({args.sp_mode}) output vs the same forward on rank 0 "
                           "alone; rel_err = max|a-b| / max|b| over the whole [1,48,32,44,80] sample")
            parity["sharded_vs_unsharded"] = cmp
        barrier()
    elif not args.no_cpu_baseline:
        parity, cpu_leg = native_block_parity(model, d_in, lat_f, h, w, dev)

    secondary = run_secondary(args, world, rank, dev, model, barrier) if not args.no_secondary else None
    if model.sequence_parallel is not None:
        model.sequence_parallel.close()  # collective: unmaps the peer buffers on every rank

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md ~1.4 PF sustained)"
    achieved_tf = attn_flops / (attn_mean_ms * 1e-3) / 1e12 if attn_mean_ms > 0 else 0.0
    # DRAM bytes of ONE launch of the kernel that was timed: the ncu capture is of the 24-head, 28160-token launch;
    # a launch over fewer heads (Ulysses) moves proportionally less (every head streams its own q/k/v/o once), and
    # a different token count has no capture -> null
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "attention_dram_traffic.json")))
        if (nq, nk, hd) == (cap["nq"], cap["nk"], cap["head_dim"]):
            traffic = int(cap["bytes_per_launch"] * hh / cap["heads"])
    except Exception:
        pass
    total_flops = flops_per_forward(tokens)
    line = {
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Wan2.2-TI2V-5B FrameINO one denoise-step forward, {args.height}x{args.width}x{args.frames} + 1 ID frame, "
                               f"{tokens} tokens, B=1, per-token timesteps, 512 text tokens",
                   "parallelism": "single GPU" if world == 1 else
                   f"ulysses sequence parallel x{world}, exchange={args.sp_mode}"
                   + (" (fused into the RoPE-prologue / attention-epilogue stores over NVLink peer memory)"
                      if args.sp_mode == "peer" else " (all_to_all_single)"),
                   "l2": "inputs larger than L2 (activations 173 MB per [N,D] tensor, weights 10 GB per forward)"},
        "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "kernel": f"attn_fwd_kernel<{hd}> (tcgen05 flash attention), {len(attn_ms)} launches, "
                               f"{nq}x{nk} tokens x {hh} heads per launch, mean {attn_mean_ms:.3f} ms",
                     "peak_source": peak_src},
        "model_tflops": total_flops / (ms * 1e-3) / 1e12 * (1.0),
        "attn_ms_per_step": sum(attn_ms) / args.steps,
        "attn_tflops": achieved_tf,
    }
    if parity:
        line["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_leg
    if secondary:
        line["secondary"] = secondary
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=121, help="pixel frames of the canvas (121 = BASELINE config 2)")
    ap.add_argument("--height", type=int, default=704, help="canvas height in pixels (multiple of 32)")
    ap.add_argument("--width", type=int, default=1280, help="canvas width in pixels (multiple of 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (cpu_baseline + parity)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary workloads (config 3 CogVideoX B=2; the Wan VAE; at N=8 config 5, 39936 tokens)")
    ap.add_argument("--cog-sp", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--no-cog-sp", action="store_true", help="N > 1: skip the sequence-parallel CogVideoX secondary")
    ap.add_argument("--cog-sp-mode", default="peer", choices=["peer", "nccl"], help="N > 1: exchange of the CogVideoX secondary")
    ap.add_argument("--sp-mode", default="peer", choices=["peer", "nccl"],
                    help="N > 1: Ulysses exchange fused over NVLink peer memory (default) or NCCL all_to_all_single")
    args = ap.parse_args()
    lat_f, h, w, tokens = workload(args.frames, args.height, args.width)
    if args.impl == "reference":
        run_reference(args, lat_f, h, w, tokens)
    else:
        run_native(args, lat_f, h, w, tokens)


if __name__ == "__main__":
    main()
