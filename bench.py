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
the timed steps, against the measured sustained bf16 peak), cpu_baseline (the CPU oracle on this box's host cores, a
bounded sample), clocks sampled during the timed region, and gpu_launches (kernels of libframeino_b200.so launched in
the timed region).
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
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
class OracleBlockSample:
    """One full-width Wan block (D 3072, FFN 14336, 24 heads) of the CPU oracle in fp32 over a 1/8 token sample
    (n_s = tokens/8 query tokens, text 512). The block is timed in two parts — everything token-linear, and the
    self-attention core — and extrapolated to the full forward with the algorithmic ratios: linear x8, attention x64,
    x30 layers. (A full forward is ~13-15 min on 8 cores; the sample keeps the run to seconds.)"""

    def __init__(self, tokens: int):
        import torch

        from frameino_b200 import synth
        from oracle import wan_oracle

        self.torch = torch
        self.wo = wan_oracle
        cfg = dict(synth.WAN22_5B)
        cfg["num_layers"] = 1
        self.cfg = wan_oracle.WanConfig(**cfg)
        shapes = {k: v for k, v in synth.wan_param_shapes(cfg).items() if k.startswith("blocks.0.")}
        self.sd = synth.make_state_dict(shapes, seed=0)
        self.tokens = tokens
        self.n_s = max(128, tokens // 8)
        g = torch.Generator().manual_seed(0)
        d = self.cfg.inner_dim
        self.x = torch.randn(1, self.n_s, d, generator=g)
        self.text = torch.randn(1, 512, d, generator=g)
        self.temb = torch.randn(1, 6, d, generator=g) * 0.1
        ang = torch.rand(self.n_s, 64, generator=g) * 6.28
        self.rot = (ang.cos().repeat_interleave(2, 1)[None, None], ang.sin().repeat_interleave(2, 1)[None, None])

    def step_ms(self) -> float:
        """Extrapolated ms per full forward from one sample."""
        torch, wo = self.torch, self.wo
        t_attn = [0.0]
        real = wo.sdpa

        def timed_sdpa(q, k, v):
            t0 = time.perf_counter()
            o = real(q, k, v)
            if q.shape[2] == k.shape[2]:  # self-attention core only
                t_attn[0] += time.perf_counter() - t0
            return o

        wo.sdpa = timed_sdpa
        try:
            t0 = time.perf_counter()
            with torch.no_grad():
                wo.wan_block(self.sd, 0, self.cfg, self.x, self.text, self.temb, self.rot)
            total = time.perf_counter() - t0
        finally:
            wo.sdpa = real
        ratio = self.tokens / self.n_s
        lin = total - t_attn[0]
        return 30.0 * (lin * ratio + t_attn[0] * ratio * ratio) * 1e3

    def describe(self) -> str:
        return (f"CPU oracle (oracle/wan_oracle.py, fp32): 1 of 30 full-width blocks over {self.n_s} of {self.tokens} "
                "tokens; token-linear time x8 + self-attention time x64, x30 layers (extrapolated)")


def run_reference(args, tokens):
    """--impl reference: the reference's CPU implementation of the path. The reference itself cannot be installed
    here (needs diffusers, absent from the image and the wheelhouse), so this times the oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use the whole host
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    sample = OracleBlockSample(tokens)
    for _ in range(args.warmup):
        sample.step_ms()
    vals = [sample.step_ms() for _ in range(args.steps)]
    v = sum(vals) / len(vals)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Wan2.2-TI2V-5B FrameINO one denoise-step forward, {args.height}x{args.width}x{args.frames} + 1 ID frame, "
                               f"{tokens} tokens, B=1 (CPU oracle port)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample.describe()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


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
        s.record()
        for _ in range(steps):
            fn()
        e.record()
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
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "attention_dram_traffic.json"))).get("bytes_per_launch")
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
    if world == 1 and not args.no_cpu_baseline:
        sample = OracleBlockSample(tokens)
        sample.step_ms()
        vals = [sample.step_ms() for _ in range(2)]
        line["cpu_baseline"] = {"value": sum(vals) / len(vals), "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "sample": sample.describe()}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sp-mode", default="peer", choices=["peer", "nccl"],
                    help="N > 1: Ulysses exchange fused over NVLink peer memory (default) or NCCL all_to_all_single")
    args = ap.parse_args()
    lat_f, h, w, tokens = workload(args.frames, args.height, args.width)
    if args.impl == "reference":
        run_reference(args, tokens)
    else:
        run_native(args, lat_f, h, w, tokens)


if __name__ == "__main__":
    main()
