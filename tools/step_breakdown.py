#!/usr/bin/env python
"""In-situ breakdown of one denoise-step forward: every ops.* call of the timed steps is bracketed with CUDA events
on the launching stream, so the durations are the ones the kernels have INSIDE a long step (sustained clocks, warm
L2) — unlike ncu's serialised cold-cache launch list. Writes a JSON summary grouped by (op, shape).

  python tools/step_breakdown.py [--steps 3] [--model wan|cog] [--out gpurun_out/step_breakdown.json]
"""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import ops, synth  # noqa: E402

OPS = ["linear", "attention", "ln_modulate", "qk_norm_rope", "gate_residual", "patchify", "unpatchify",
       "linear_small_m", "build_mod_table", "timestep_embedding", "swap01"]


def shape_key(name, args, kwargs):
    def sh(t):
        return "x".join(str(s) for s in t.shape) if isinstance(t, torch.Tensor) else "-"

    if name == "linear":
        x, w = args[0], args[1]
        return f"M{x.numel() // x.shape[-1]} N{w.shape[0]} K{w.shape[1]} epi{kwargs.get('epilogue', 0)}"
    if name in ("attention", "sp_attention", "sp_attention_scatter"):
        q, k = args[0], args[1]
        return f"q{sh(q)} k{sh(k)} h{args[3] if len(args) > 3 else kwargs.get('heads')}"
    if name in ("peer_barrier", "all_to_all_single"):
        return "-"
    return sh(args[0]) if args else "-"


def flops(name, args, kwargs):
    if name == "linear":
        x, w = args[0], args[1]
        return 2.0 * (x.numel() // x.shape[-1]) * w.shape[0] * w.shape[1]
    if name in ("attention", "sp_attention", "sp_attention_scatter"):
        q, k = args[0], args[1]
        return 4.0 * q.shape[0] * q.shape[1] * k.shape[1] * q.shape[2]
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--model", default="wan", choices=["wan", "cog"])
    ap.add_argument("--out", default="gpurun_out/step_breakdown.json")
    ap.add_argument("--sp-check", action="store_true", help="N > 1: run tools/sp_check.py's parity cases first")
    ap.add_argument("--sp-mode", default="peer", choices=["peer", "nccl"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:  # torchrun: Ulysses sequence parallel, rank 0 reports its own breakdown
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        if args.sp_check:
            import sp_check

            sp_check.run_cases()
    if args.model == "wan":
        cfg = synth.WAN22_5B
        model = synth.build_wan_on_device(cfg, seed=0, device=dev)
        hidden, ts, text = synth.make_wan_inputs(cfg, 31, 44, 80, n_id=1, text_len=512, text_true_len=120,
                                                 dtype=torch.bfloat16)
        inputs = dict(hidden_states=hidden.to(dev), timestep=ts.to(dev), encoder_hidden_states=text.to(dev),
                      return_dict=False)
    else:
        cfg = synth.COG_5B_I2V
        model = synth.build_cog_on_device(cfg, seed=0, device=dev)
        lat_f, h, w = 13, 60, 90
        hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=1, dtype=torch.bfloat16)
        cos, sin = synth.cog_rope_tables(64, h // 2, w // 2, lat_f, 1, device=dev)
        inputs = dict(hidden_states=hidden.to(dev), encoder_hidden_states=text.to(dev), timestep=ts.to(dev),
                      image_rotary_emb=(cos, sin), return_dict=False)
    sp = None
    if world > 1:
        from frameino_b200.ulysses import enable_sequence_parallel

        sp = enable_sequence_parallel(model, mode=args.sp_mode)
    for _ in range(2):
        model(**inputs)
    torch.cuda.synchronize()

    records = []
    real = {n: getattr(ops, n) for n in OPS if hasattr(ops, n)}
    sp_real = {}
    if sp is not None:
        sp_real = {n: getattr(sp, n) for n in ("_qkv_scatter", "_attention_scatter", "_swap01", "_attention")}

    def wrap(name, fn):
        def inner(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            records.append((name, shape_key(name, a, kw), flops(name, a, kw), s, e))
            return r

        return inner

    for n, fn in real.items():
        setattr(ops, n, wrap(n, fn))
    for n, fn in sp_real.items():
        setattr(sp, n, wrap("sp" + n, fn))
    if sp is not None:
        real_barrier = ops.peer_barrier
        ops.peer_barrier = wrap("peer_barrier", real_barrier)
        real_a2a = dist.all_to_all_single
        dist.all_to_all_single = wrap("all_to_all_single", real_a2a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        model(**inputs)
    e.record()
    torch.cuda.synchronize()
    for n, fn in real.items():
        setattr(ops, n, fn)
    for n, fn in sp_real.items():
        setattr(sp, n, fn)
    if sp is not None:
        ops.peer_barrier = real_barrier
        dist.all_to_all_single = real_a2a
        sp.close()
    if rank != 0:
        dist.destroy_process_group()
        return
    total = s.elapsed_time(e) / args.steps
    agg = collections.OrderedDict()
    for name, key, fl, a, b in records:
        d = agg.setdefault((name, key), {"n": 0, "ms": 0.0, "flops": 0.0})
        d["n"] += 1
        d["ms"] += a.elapsed_time(b)
        d["flops"] += fl
    rows = []
    for (name, key), d in agg.items():
        rows.append({"op": name, "shape": key, "calls_per_step": d["n"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                     "avg_us": 1e3 * d["ms"] / d["n"],
                     "tflops": (d["flops"] / (d["ms"] * 1e-3) / 1e12) if d["flops"] and d["ms"] > 0 else None})
    rows.sort(key=lambda r: -r["ms_per_step"])
    covered = sum(r["ms_per_step"] for r in rows)
    out = {"model": args.model, "world": world, "steps": args.steps, "ms_per_step": total, "ms_in_ops": covered, "rows": rows}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(f"step {total:.2f} ms, inside ops {covered:.2f} ms")
    for r in rows:
        tf = f"{r['tflops']:8.1f} TF/s" if r["tflops"] else ""
        print(f"{r['ms_per_step']:9.3f} ms  n={r['calls_per_step']:5.1f}  avg {r['avg_us']:9.1f} us  {r['op']:<18} {r['shape']} {tf}")


if __name__ == "__main__":
    main()
