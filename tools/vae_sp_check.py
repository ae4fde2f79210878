#!/usr/bin/env python
"""Multi-GPU check of the row-parallel VAE decode and encode (frameino_b200/vae.py RowParallel) against the un-sharded
ones of the same model on the same inputs — expected bit-identical — and their timing at config 2's canvas.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/vae_sp_check.py [--full] [--out f.json]

Cases: the tiny VAE at three canvases (uneven bands, one latent row per rank), the real widths at a small canvas, and
with --full the 704x1280x121 decode (un-sharded on every rank first, then sharded; max over ranks of the device time)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import synth  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    y = fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return y, float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"n_gpus": world, "cases": []}
    ok = True
    cases = [("tiny", synth.VAE_TINY, 3, max(4, world), 6), ("tiny", synth.VAE_TINY, 2, world + 3, 8),
             ("tiny", synth.VAE_TINY, 1, 2 * world + 1, 8), ("real widths", synth.WAN22_VAE, 2, max(6, world), 12)]
    for name, cfg, tl, h, w in cases:
        vae = synth.build_vae_on_device(cfg, seed=1, device=dev)
        g = torch.Generator(device=dev).manual_seed(5)
        z = torch.randn(1, cfg["z_dim"], tl, h, w, generator=g, device=dev)
        ref = vae.decode(z, return_dict=False)[0]
        vae.enable_row_parallel()
        got = vae.decode(z, return_dict=False)[0]
        clip = ref[:, :, : 1 + 4 * ((ref.shape[2] - 1) // 4)].contiguous()
        enc_rp = vae.encode(clip).latent_dist.parameters
        vae.disable_row_parallel()
        enc = vae.encode(clip).latent_dist.parameters
        ediff = (enc_rp - enc).abs().max().reshape(1)
        dist.all_reduce(ediff, op=dist.ReduceOp.MAX)
        diff = float((got - ref).abs().max())
        t = torch.tensor([diff], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        case = {"vae": name, "latent": [tl, h, w], "max_abs_diff_vs_unsharded": float(t.item()),
                "equal": bool(t.item() == 0.0), "shape": list(got.shape),
                "encode_max_abs_diff_vs_unsharded": float(ediff.item()), "encode_shape": list(enc_rp.shape)}
        ok = ok and got.shape == ref.shape and enc_rp.shape == enc.shape and case["max_abs_diff_vs_unsharded"] <= 1e-2 \
            and case["encode_max_abs_diff_vs_unsharded"] <= 1e-2
        res["cases"].append(case)
        del vae
        torch.cuda.empty_cache()
    # fewer latent rows than ranks: falls back to whole frames on every rank (same result)
    if world > 2:
        vae = synth.build_vae_on_device(synth.VAE_TINY, seed=1, device=dev)
        z = torch.randn(1, 16, 2, 2, 8, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
        ref = vae.decode(z, return_dict=False)[0]
        vae.enable_row_parallel()
        ok = ok and torch.equal(vae.decode(z, return_dict=False)[0], ref)
        res["fallback_fewer_rows_than_ranks"] = bool(ok)
        del vae
    if args.full:
        cfg = synth.WAN22_VAE
        vae = synth.build_vae_on_device(cfg, seed=0, device=dev)
        g = torch.Generator(device=dev).manual_seed(3)
        z = torch.randn(1, 48, 31, 44, 80, generator=g, device=dev)
        vae.decode(z[:, :, :2], return_dict=False)
        ref, ms_1 = timed(lambda: vae.decode(z, return_dict=False)[0])
        vae.enable_row_parallel()
        vae.decode(z[:, :, :2], return_dict=False)
        got, ms_n = timed(lambda: vae.decode(z, return_dict=False)[0])
        diff = (got - ref).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        res["full_704x1280x121"] = {"unsharded_ms": ms_1, "row_parallel_ms": ms_n, "speedup": ms_1 / ms_n,
                                    "max_abs_diff_vs_unsharded": float(diff.item()), "finite": bool(torch.isfinite(got).all())}
        ok = ok and float(diff.item()) <= 1e-2
        del got, ref
        x = torch.randn(1, 3, 121, 704, 1280, generator=g, device=dev).clamp_(-1, 1)
        vae.encode(x[:, :, :5])
        enc_n, ems_n = timed(lambda: vae.encode(x).latent_dist.parameters)
        vae.disable_row_parallel()
        vae.encode(x[:, :, :5])
        enc_1, ems_1 = timed(lambda: vae.encode(x).latent_dist.parameters)
        ediff = (enc_n - enc_1).abs().max().reshape(1)
        dist.all_reduce(ediff, op=dist.ReduceOp.MAX)
        res["full_704x1280x121_encode"] = {"unsharded_ms": ems_1, "row_parallel_ms": ems_n, "speedup": ems_1 / ems_n,
                                           "max_abs_diff_vs_unsharded": float(ediff.item())}
        ok = ok and float(ediff.item()) <= 1e-2
    res["ok"] = bool(ok)
    if rank == 0:
        print("VAE_SP " + json.dumps(res))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
