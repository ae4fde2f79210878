#!/usr/bin/env python
"""Tensor-for-tensor comparison of the native models with the REAL reference stack (the reference's own
architecture/*.py on top of an installed `diffusers`), for a machine that has both — this build container and the GPU
boxes do not (no diffusers, no network), so the committed parity evidence is the oracle + the golden vectors generated
through tests/golden/diffusers_shim; this script is what closes the "parity unpinned" gap for the upstream-only pieces
(FeedForward, FP32LayerNorm, RMSNorm, AdaLayerNorm, CogVideoXLayerNormZero) wherever diffusers exists:

  pip install git+https://github.com/huggingface/diffusers.git          # what the reference's requirements.txt:12 asks for
  python tools/compare_with_diffusers.py --reference /path/to/FrameINO [--config small|tiny|5b] [--model wan|cog]

Both sides load the same seeded synthetic state dict (diffusers key names are kept by the native modules), run the same
seeded inputs on cuda:0 in bf16, and the report lists per-block max|a-b|/max|b| and the final cosine; the bar is the
north star's (2e-2 per layer, cosine >= 0.999). Exit code 1 if it is missed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frameino_b200 import synth  # noqa: E402


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def cosine(a, b):
    return float(torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0))


def run_wan(ref_root: str, cfg_name: str):
    from architecture.transformer_wan import WanTransformer3DModel as RefWan  # the reference class on real diffusers

    from frameino_b200.wan import WanTransformer3DModel as FinoWan

    cfg, shape = {"tiny": (synth.WAN_TINY, (5, 16, 16)), "small": (synth.WAN_SMALL, (3, 32, 32)),
                  "5b": (synth.WAN22_5B, (31, 44, 80))}[cfg_name]
    dev = torch.device("cuda", 0)
    if cfg_name == "5b":
        fino = synth.build_wan_on_device(cfg, seed=0, device=dev)
        sd = fino.state_dict()
        ref = RefWan(**cfg)
        ref.load_state_dict({k: v.cpu() for k, v in sd.items()})
    else:
        sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
        fino = FinoWan(**cfg)
        fino.load_state_dict(sd)
        fino = fino.to_inference_dtype(torch.bfloat16).to(dev).eval()
        ref = RefWan(**cfg)
        ref.load_state_dict(sd)
    # diffusers' from_pretrained(torch_dtype=bf16) policy: bf16 except _keep_in_fp32_modules (transformer_wan.py:393)
    keep = getattr(RefWan, "_keep_in_fp32_modules", None) or []
    for name, p in ref.named_parameters():
        p.data = p.data.to(torch.float32 if any(k in name for k in keep) else torch.bfloat16)
    ref = ref.to(dev).eval()
    hidden, ts, text = synth.make_wan_inputs(cfg, *shape, n_id=1, text_len=64, text_true_len=40, dtype=torch.bfloat16)
    args = dict(hidden_states=hidden.to(dev), timestep=ts.to(dev), encoder_hidden_states=text.to(dev), return_dict=False)
    ref_taps, taps = {}, {}
    hooks = [blk.register_forward_hook(lambda m, a, o, i=i: ref_taps.__setitem__(f"blocks.{i}.out", o.detach().clone()))
             for i, blk in enumerate(ref.blocks)]
    fino.__dict__["_fino_taps"] = taps
    with torch.no_grad():
        y_ref = ref(**args)[0]
        y = fino(**args)[0]
    for h in hooks:
        h.remove()
    rows = {k: rel_err(taps[k], v) for k, v in ref_taps.items()}
    return rows, rel_err(y, y_ref), cosine(y, y_ref)


def run_cog(ref_root: str, cfg_name: str):
    from architecture.cogvideox_transformer_3d import CogVideoXTransformer3DModel as RefCog
    from architecture.embeddings import get_3d_rotary_pos_embed

    from frameino_b200.cogvideox import CogVideoXTransformer3DModel as FinoCog

    cfg, (f, h, w) = {"tiny": (synth.COG_TINY, (3, 12, 16)), "small": (synth.COG_TINY, (5, 24, 32)),
                      "5b": (synth.COG_5B_I2V, (13, 60, 90))}[cfg_name]
    dev = torch.device("cuda", 0)
    sd = synth.make_cog_state_dict(cfg, seed=0, dtype=torch.bfloat16) if cfg_name != "5b" else None
    if sd is None:
        fino = synth.build_cog_on_device(cfg, seed=0, device=dev)
        sd = {k: v.cpu() for k, v in fino.state_dict().items()}
    else:
        fino = FinoCog(**cfg)
        fino.load_state_dict(sd)
        fino = fino.to(torch.bfloat16).to(dev).eval()
    ref = RefCog(**cfg)
    ref.load_state_dict(sd)
    ref = ref.to(torch.bfloat16).to(dev).eval()
    hidden, ts, text = synth.make_cog_inputs(cfg, f, h, w, n_id=1, batch=2, dtype=torch.bfloat16)
    p = cfg.get("patch_size", 2)
    cos, sin = get_3d_rotary_pos_embed(cfg["attention_head_dim"], ((0, 0), (h // p, w // p)), (h // p, w // p), f,
                                       device=dev)
    # FrameINO appends frame 0's table for the ID frame (pipeline_cogvideox_i2v_motion_FrameINO.py:834-839)
    per = (h // p) * (w // p)
    cos, sin = torch.cat([cos, cos[:per]]), torch.cat([sin, sin[:per]])
    args = dict(hidden_states=hidden.to(dev), encoder_hidden_states=text.to(dev), timestep=ts.to(dev),
                image_rotary_emb=(cos, sin), return_dict=False)
    ref_taps, taps = {}, {}
    hooks = [blk.register_forward_hook(
        lambda m, a, o, i=i: ref_taps.__setitem__(f"transformer_blocks.{i}.out", o[0].detach().clone()))
        for i, blk in enumerate(ref.transformer_blocks)]
    fino.__dict__["_fino_taps"] = taps
    with torch.no_grad():
        y_ref = ref(**args)[0]
        y = fino(**args)[0]
    for hk in hooks:
        hk.remove()
    rows = {k: rel_err(taps[k], v) for k, v in ref_taps.items()}
    return rows, rel_err(y, y_ref), cosine(y, y_ref)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of UVA-Computer-Vision-Lab/FrameINO")
    ap.add_argument("--model", default="wan", choices=["wan", "cog"])
    ap.add_argument("--config", default="small", choices=["tiny", "small", "5b"])
    args = ap.parse_args()
    try:
        import diffusers  # noqa: F401
    except ImportError:
        sys.exit("this comparison needs the real `diffusers` package (see the module docstring); "
                 "without it the parity evidence is tests/test_oracle_golden.py + the -m gpu suite")
    os.chdir(args.reference)  # the reference appends abspath('.') to sys.path for its local imports
    sys.path.insert(0, args.reference)
    rows, err, cos = (run_wan if args.model == "wan" else run_cog)(args.reference, args.config)
    worst = max(rows.values()) if rows else 0.0
    print(json.dumps({"model": args.model, "config": args.config, "per_block_rel_err": rows, "worst_block": worst,
                      "final_rel_err": err, "final_cosine": cos}, indent=1))
    sys.exit(0 if (worst <= 2e-2 and cos >= 0.999) else 1)


if __name__ == "__main__":
    main()
