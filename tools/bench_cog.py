#!/usr/bin/env python
"""Secondary workload (BASELINE.json configs[2]): CogVideoX-5B-I2V FrameINO one denoise step, 480x720x49 + 1 ID frame
(S = 226 text + 17550 video + 1350 ID = 19126 tokens), batch 2 (the pipeline's batched CFG) or 1, bf16, 1 x B200.
Prints one JSON line in the bench.py format (not the headline metric; bench.py measures the Wan config)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from frameino_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = synth.COG_5B_I2V
    model = synth.build_cog_on_device(cfg, seed=0, device=dev)
    lat_f, h, w = 13, 60, 90
    hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=args.batch, dtype=torch.bfloat16)
    cos, sin = synth.cog_rope_tables(64, h // 2, w // 2, lat_f, 1, device=dev)
    d_in = [hidden.to(dev), text.to(dev), ts.to(dev)]
    seq = 226 + (lat_f + 1) * (h // 2) * (w // 2)

    def step():
        return model(hidden_states=d_in[0], encoder_hidden_states=d_in[1], timestep=d_in[2], image_rotary_emb=(cos, sin),
                     return_dict=False)[0]

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    with bench.ClockSampler(0) as clocks:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        s.record()
        for _ in range(args.steps):
            step()
        e.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    d, f, layers = 3072, 12288, 42
    flops = args.batch * layers * (3 * 2 * seq * d * d + 4 * seq * seq * d + 2 * seq * d * d + 4 * seq * d * f)
    print(json.dumps({
        "metric": "cogvideox_5b_i2v_frameino_denoise_step_ms", "value": ms, "unit": "ms", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"CogVideoX-5B-I2V FrameINO one denoise step, 480x720x49 + 1 ID frame, S={seq}, B={args.batch}"},
        "gpu_launches": ops.launch_count() - l0, "clocks": clocks.summary(), "model_tflops": flops / (ms * 1e-3) / 1e12,
    }))


if __name__ == "__main__":
    main()
