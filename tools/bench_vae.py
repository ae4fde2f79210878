#!/usr/bin/env python
"""Wan2.2 VAE at the BASELINE config-2 canvas: decode of the [1, 48, 31, 44, 80] latent to
704x1280x121 and encode of a 704x1280x121 clip, random-init weights of the TI2V-5B VAE architecture (704.7 M
parameters). Prints one JSON line: ms, algorithmic TFLOP (convolutions + attention), TFLOP/s.
Usage: python tools/bench_vae.py [--frames 121] [--height 704] [--width 1280] [--what decode encode] [--breakdown]"""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import ops, synth  # noqa: E402


def build(cfg, dev):
    return synth.build_vae_on_device(cfg, seed=0, device=dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=121)
    ap.add_argument("--height", type=int, default=704)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--what", nargs="+", default=["decode", "encode"])
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = synth.WAN22_VAE
    vae = build(cfg, dev)
    tl, h, w = (args.frames - 1) // 4 + 1, args.height // 16, args.width // 16
    res = {"workload": f"Wan2.2-TI2V-5B VAE, {args.height}x{args.width}x{args.frames} (latent {tl}x{h}x{w}x48), bf16 "
                       "activations, fp32 accumulation", "params_m": sum(p.numel() for p in vae.parameters()) / 1e6}
    flops = [0.0]
    times = collections.defaultdict(lambda: [0.0, 0, 0.0])
    events = []
    real_conv, real_linear = ops.conv3d_cl, ops.linear

    def conv(x, weight, bias, kernel, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.breakdown:
            s.record()
        y = real_conv(x, weight, bias, kernel, **kw)
        f = 2.0 * y.shape[0] * y.shape[1] * y.shape[2] * weight.shape[0] * kernel[0] * kernel[1] * kernel[2] * x.shape[3]
        flops[0] += f
        if args.breakdown:
            e.record()
            events.append((f"conv k{kernel} Cin{x.shape[3]} Cout{weight.shape[0]} {y.shape[0]}x{y.shape[1]}x{y.shape[2]}", s, e, f))
        return y

    def linear(x, weight, bias=None, **kw):
        y = real_linear(x, weight, bias, **kw)
        flops[0] += 2.0 * (x.numel() // x.shape[-1]) * weight.shape[0] * weight.shape[1]
        return y

    ops.conv3d_cl, ops.linear = conv, linear
    if args.breakdown:  # every other op of the path, by name (CUDA events on the launching stream)
        from frameino_b200 import vae as vae_mod

        def timed(name, fn):
            def w(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*a, **k)
                e.record()
                events.append((name, s, e, 0.0))
                return r
            return w

        for name in ("rms_act_cl", "upsample2x_cl", "dupup_add_cl", "avgdown_add_cl", "softmax_rows", "vae_to_cl", "vae_from_cl"):
            setattr(ops, name, timed(name, getattr(ops, name)))
        ops.linear = timed("linear (1x1 convs, attention GEMMs)", ops.linear)
        vae_mod._ConvCaches.advance = timed("cache advance (copy 2 frames)", vae_mod._ConvCaches.advance)
        vae_mod._ConvCaches.input = timed("cache input view (alloc + zero on first use)", vae_mod._ConvCaches.input)
    for what in args.what:
        if what == "decode":
            inp = torch.randn(1, 48, tl, h, w, device=dev)
            fn = lambda: vae.decode(inp, return_dict=False, output_dtype=torch.bfloat16)[0]  # noqa: E731
        else:
            inp = (torch.rand(1, 3, args.frames, args.height, args.width, device=dev) * 2 - 1).bfloat16()
            fn = lambda: vae.encode(inp).latent_dist.mode()  # noqa: E731
        out = fn()  # warm-up
        torch.cuda.synchronize()
        flops[0] = 0.0
        events.clear()
        l0 = ops.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        res[what] = {"ms": ms, "tflop": flops[0] / 1e12, "tflops": flops[0] / ms / 1e9, "launches": ops.launch_count() - l0,
                     "finite": bool(torch.isfinite(out.float()).all()), "out_shape": list(out.shape),
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        if args.breakdown:
            agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
            for name, s_, e_, f in events:
                a = agg[name]
                a[0] += s_.elapsed_time(e_)
                a[1] += 1
                a[2] += f
            rows = [{"op": k, "calls": v[1], "ms": round(v[0], 2), "tflops": round(v[2] / v[0] / 1e9, 1) if v[2] else None}
                    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])]
            res[what]["conv_ms"] = round(sum(r["ms"] for r in rows if r["op"].startswith("conv k")), 1)
            res[what]["other_ms"] = {r["op"]: r["ms"] for r in rows if not r["op"].startswith("conv k")}
            res[what]["top_convs"] = [r for r in rows if r["op"].startswith("conv k")][:14]
        del out, inp
        torch.cuda.empty_cache()
    print("VAE " + json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
