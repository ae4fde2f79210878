#!/usr/bin/env python
"""The whole Wan FrameINO image-to-video call, pixels to pixels, on the device (frameino_b200.pipeline.WanFrameINOPipeline
= reference pipelines/pipeline_wan_i2v_motion_FrameINO.py __call__): 3 VAE encodes (first-frame canvas, trajectory
video, ID image) -> 50 scheduler steps x 2 CFG forwards of the Wan2.2-5B DiT -> VAE decode, at BASELINE config 2
(704x1280x121), random-init weights of both architectures, synthetic pixels and prompt embeddings.

  python tools/run_pipeline.py [--steps 50] [--out profiles/rNN_pipeline_e2e.json]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_pipeline.py   (Ulysses, VAE replicated)

Prints one JSON line with the stage times (CUDA events on the launching stream; host inputs, host output: the H2D copies
of the pixels and the D2H copy of the video are inside the total)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import synth  # noqa: E402
from frameino_b200.pipeline import WanFrameINOPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--frames", type=int, default=121)
    ap.add_argument("--height", type=int, default=704)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-vae-parallel", action="store_true", help="keep the VAE decode replicated under torchrun")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vcfg = synth.with_latent_stats(synth.WAN22_VAE)
    vae = synth.build_vae_on_device(vcfg, seed=1, device=dev)
    tf = synth.build_wan_on_device(synth.WAN22_5B, seed=0, device=dev)
    if world > 1:
        from frameino_b200.ulysses import enable_sequence_parallel

        enable_sequence_parallel(tf)
        if not args.no_vae_parallel:
            vae.enable_row_parallel()  # encodes and decode split by frame rows
    pipe = WanFrameINOPipeline(vae=vae, transformer=tf)
    inp = synth.make_pipeline_inputs(vcfg, synth.WAN22_5B["text_dim"], args.frames, args.height, args.width, n_id=1,
                                     text_len=512)
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    host["prompt_embeds"] = host["prompt_embeds"].bfloat16()
    host["negative_prompt_embeds"] = host["negative_prompt_embeds"].bfloat16()
    stages = {}
    real_prepare, real_decode = pipe.prepare_latents, vae.decode

    def timed(name, fn):
        def w(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            stages[name] = (s, e)
            return r
        return w

    pipe.prepare_latents = timed("prepare_latents_3_vae_encodes", real_prepare)
    vae.decode = timed("vae_decode", real_decode)

    def call(steps):
        v = pipe(image=host["image"], traj_tensor=host["traj_tensor"], ID_tensor=host["ID_tensor"],
                 prompt_embeds=host["prompt_embeds"], negative_prompt_embeds=host["negative_prompt_embeds"],
                 latents=host["latents"], height=args.height, width=args.width, num_frames=args.frames,
                 num_inference_steps=steps, guidance_scale=5.0, output_type="pt").frames
        return v.cpu() if rank == 0 else v  # every rank holds the video; rank 0 hands it to the host

    call(1)  # warm-up: weight packs, workspaces, peer buffers
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    video = call(args.steps)
    e.record()
    torch.cuda.synchronize()
    total = s.elapsed_time(e)
    t = torch.tensor([total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"workload": f"Wan FrameINO pipeline call, {args.height}x{args.width}x{args.frames}, {args.steps} steps x 2 "
                       f"CFG forwards, guidance 5.0, random-init Wan2.2-5B + Wan2.2 VAE, synthetic inputs",
           "n_gpus": world, "parallelism": "single GPU" if world == 1 else
           f"DiT: ulysses x{world}; VAE encode / decode: " + ("replicated" if args.no_vae_parallel else f"row bands x{world}"), "total_ms": float(t.item()),
           "stage_ms": {k: a.elapsed_time(b) for k, (a, b) in stages.items()},
           "video_shape": list(video.shape), "finite": bool(torch.isfinite(video).all()),
           "h2d_bytes": int(sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))),
           "d2h_bytes": int(video.numel() * video.element_size())}
    res["stage_ms"]["denoise_loop_postprocess_and_copies"] = res["total_ms"] - sum(res["stage_ms"].values())
    if rank == 0:
        print("PIPELINE " + json.dumps(res))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
