#!/usr/bin/env python
"""BASELINE.json configs[3]: the full FrameINO Wan2.2-5B sampling loop (50 flow-match Euler steps x 2 CFG forwards,
guidance 5.0) with Ulysses sequence parallelism. Launch with torchrun (or plain python for 1 GPU):

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_loop.py --steps 50 [--compare]

Prints the wall/device time of the loop and, with --compare, the cosine between the final latents of the
sequence-parallel run and of the same loop on one GPU (rank 0 runs it un-sharded afterwards)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import synth  # noqa: E402
from frameino_b200.sampling import wan_frameino_denoise, wan_frameino_denoise_fused  # noqa: E402
from frameino_b200.ulysses import disable_sequence_parallel, enable_cfg_parallel, enable_sequence_parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--frames", type=int, default=121)
    ap.add_argument("--height", type=int, default=704)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--compare", action="store_true")
    ap.add_argument("--fused", action="store_true", help="device-side loop glue (wan_frameino_denoise_fused)")
    ap.add_argument("--cfg-parallel", action="store_true",
                    help="half of the ranks run the conditional forward, half the unconditional one (Ulysses inside each half)")
    ap.add_argument("--both", action="store_true", help="time the plain and the fused loop, report both + max diff")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.WAN22_5B
    model = synth.build_wan_on_device(cfg, seed=0, device=dev)
    lat_f, h, w = (args.frames - 1) // 4 + 1, args.height // 16, args.width // 16
    g = torch.Generator().manual_seed(11)
    c = cfg["out_channels"]
    lat = torch.randn(1, c, lat_f, h, w, generator=g)
    cond = torch.zeros(1, c, lat_f, h, w)
    cond[:, :, 0] = torch.randn(1, c, h, w, generator=g)
    mask = torch.ones(1, 1, lat_f, h, w)  # the reference's [1,1,F,H,W] first_frame_mask (pipeline :529-532)
    mask[:, :, 0] = 0
    traj = torch.randn(1, c, lat_f + 1, h, w, generator=g)
    traj[:, :, lat_f:] = 0
    idl = torch.randn(1, c, 1, h, w, generator=g)
    pos = torch.randn(1, 512, cfg["text_dim"], generator=g)
    pos[:, 120:] = 0
    neg = torch.zeros(1, 512, cfg["text_dim"])
    tensors = [t.to(dev) for t in (lat, cond, mask, traj, idl)] + [pos.to(dev).bfloat16(), neg.to(dev).bfloat16()]

    loop = wan_frameino_denoise_fused if args.fused else wan_frameino_denoise

    extra = {}

    def run(loop=loop):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = loop(model, *tensors, num_steps=args.steps, **extra)
        e.record()
        torch.cuda.synchronize()
        return out, s.elapsed_time(e)

    if world > 1 and args.cfg_parallel:
        extra["cfg_parallel"] = enable_cfg_parallel(model)
    elif world > 1:
        enable_sequence_parallel(model)
    loop(model, *tensors, num_steps=1, **extra)  # warm-up
    out_sp, ms_sp = run()
    res = {"n_gpus": world, "steps": args.steps, "forwards": 2 * args.steps, "loop": loop.__name__,
           "parallelism": ("cfg x2, ulysses x%d" % (world // 2)) if (args.cfg_parallel and world > 1) else "ulysses x%d" % world,
           "ms_per_scheduler_step": ms_sp / args.steps, "loop_ms": ms_sp,
           "ms_per_forward": ms_sp / (2 * args.steps), "tokens": (lat_f + 1) * (h // 2) * (w // 2),
           "finite": bool(torch.isfinite(out_sp).all())}
    if args.both:
        other = wan_frameino_denoise if args.fused else wan_frameino_denoise_fused
        other(model, *tensors, num_steps=1, **extra)
        out_o, ms_o = run(other)
        res[other.__name__ + "_loop_ms"] = ms_o
        res["max_abs_diff_between_loops"] = float((out_o - out_sp).abs().max())
    if args.compare and world > 1:
        disable_sequence_parallel(model)
        extra.clear()
        if rank == 0:
            out_1, ms_1 = run_single(model, tensors, args.steps)
            res["single_gpu_loop_ms"] = ms_1
            res["final_latent_cosine_vs_single_gpu"] = float(torch.nn.functional.cosine_similarity(
                out_sp.flatten().float(), out_1.flatten().float(), dim=0))
        dist.barrier()
    if rank == 0:
        print("LOOP " + json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def run_single(model, tensors, steps):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = wan_frameino_denoise(model, *tensors, num_steps=steps)
    e.record()
    torch.cuda.synchronize()
    return out, s.elapsed_time(e)


if __name__ == "__main__":
    main()
