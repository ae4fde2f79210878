#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(/root/reference/architecture/{transformer_wan,cogvideox_transformer_3d,attention_processor,embeddings}.py) on CPU in
fp32, with tests/golden/diffusers_shim standing in for the absent `diffusers` package (see its __init__ docstring for
what is restated there). Weights and inputs come from the seeded recipe in frameino_b200/synth.py, so the fixtures
hold only outputs and a few per-layer taps.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The GPU box never runs this; tests only read the committed *.pt files.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FRAMEINO_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "diffusers_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from frameino_b200 import synth  # noqa: E402


def _tap_summary(t: torch.Tensor) -> torch.Tensor:
    """First 256 values + mean |x| + max |x| of a tapped tensor (keeps fixtures small)."""
    f = t.detach().float().reshape(-1)
    return torch.cat([f[:256], f.abs().mean()[None], f.abs().max()[None]])


def wan_golden():
    os.chdir(REF)  # the reference appends abspath('.') to sys.path for its local imports
    from architecture.transformer_wan import WanTransformer3DModel  # the reference class itself

    out = {}
    for name, cfg, shape in (("tiny", synth.WAN_TINY, (5, 16, 16)), ("small", synth.WAN_SMALL, (3, 16, 16))):
        torch.manual_seed(0)
        model = WanTransformer3DModel(**cfg).eval()
        sd = synth.make_wan_state_dict(cfg, seed=0)
        missing, unexpected = model.load_state_dict(sd, strict=True)
        taps = {}
        hooks = []
        for i, blk in enumerate(model.blocks):
            hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.out", _tap_summary(o))))
            hooks.append(blk.attn1.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.attn1", _tap_summary(o))))
            hooks.append(blk.attn2.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.attn2", _tap_summary(o))))
        f, h, w = shape
        for mode in ("per_token", "scalar"):
            hidden, ts, text = synth.make_wan_inputs(cfg, f, h, w, n_id=1, text_len=16, text_true_len=11,
                                                     per_token_timestep=(mode == "per_token"))
            with torch.no_grad():
                y = model(hidden_states=hidden, timestep=ts, encoder_hidden_states=text, return_dict=False)[0]
            out[f"{name}.{mode}.sample"] = y.clone()
            for k, v in taps.items():
                out[f"{name}.{mode}.{k}"] = v.clone()
        for hk in hooks:
            hk.remove()
        # rope tables the reference module produces for this latent shape
        cos, sin = model.rope(hidden)
        out[f"{name}.rope_cos"] = cos.clone()
        out[f"{name}.rope_sin"] = sin.clone()
    torch.save(out, os.path.join(HERE, "wan_golden.pt"))
    print("wrote wan_golden.pt:", {k: tuple(v.shape) for k, v in out.items() if k.endswith("sample")})


def cog_golden():
    os.chdir(REF)
    from architecture.cogvideox_transformer_3d import CogVideoXTransformer3DModel
    from architecture.embeddings import get_3d_rotary_pos_embed

    cfg = synth.COG_TINY
    torch.manual_seed(0)
    model = CogVideoXTransformer3DModel(**cfg).eval()
    sd = synth.make_cog_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    out = {}
    lat_f = (cfg["sample_frames"] - 1) // cfg["temporal_compression_ratio"] + 1
    h, w = cfg["sample_height"], cfg["sample_width"]
    hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=2)
    gh, gw = h // cfg["patch_size"], w // cfg["patch_size"]
    # pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:540-584 (1.0 checkpoints) + :834-839 (ID frame = frame-0 rows)
    cos, sin = get_3d_rotary_pos_embed(cfg["attention_head_dim"], ((0, 0), (gh, gw)), (gh, gw), lat_f)
    cos = torch.cat([cos, cos[: gh * gw]], dim=0)
    sin = torch.cat([sin, sin[: gh * gw]], dim=0)
    taps = {}
    hooks = [blk.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.out", _tap_summary(o[0])))
             for i, blk in enumerate(model.transformer_blocks)]
    with torch.no_grad():
        y = model(hidden_states=hidden, encoder_hidden_states=text, timestep=ts, image_rotary_emb=(cos, sin),
                  return_dict=False)[0]
    out["tiny.sample"] = y.clone()
    out["tiny.rope_cos"] = cos.clone()
    out["tiny.rope_sin"] = sin.clone()
    for k, v in taps.items():
        out[f"tiny.{k}"] = v.clone()
    for hk in hooks:
        hk.remove()
    # fused-projection path of the reference (fuse_qkv_projections -> FusedCogVideoXAttnProcessor2_0)
    model.fuse_qkv_projections()
    with torch.no_grad():
        y2 = model(hidden_states=hidden, encoder_hidden_states=text, timestep=ts, image_rotary_emb=(cos, sin),
                   return_dict=False)[0]
    out["tiny.sample_fused"] = y2.clone()
    # non-default canvas: trilinear resize of the positional table (embeddings.py:782-798)
    hidden_b, ts_b, text_b = synth.make_cog_inputs(cfg, lat_f, h + 4, w - 4, n_id=1, batch=1, seed=3)
    gh2, gw2 = (h + 4) // cfg["patch_size"], (w - 4) // cfg["patch_size"]
    cos2, sin2 = get_3d_rotary_pos_embed(cfg["attention_head_dim"], ((0, 0), (gh2, gw2)), (gh2, gw2), lat_f)
    cos2 = torch.cat([cos2, cos2[: gh2 * gw2]], dim=0)
    sin2 = torch.cat([sin2, sin2[: gh2 * gw2]], dim=0)
    model.unfuse_qkv_projections()
    with torch.no_grad():
        y3 = model(hidden_states=hidden_b, encoder_hidden_states=text_b, timestep=ts_b, image_rotary_emb=(cos2, sin2),
                   return_dict=False)[0]
    out["tiny.sample_resized"] = y3.clone()
    out["tiny.rope_cos_resized"] = cos2.clone()
    out["tiny.rope_sin_resized"] = sin2.clone()
    torch.save(out, os.path.join(HERE, "cog_golden.pt"))
    print("wrote cog_golden.pt:", {k: tuple(v.shape) for k, v in out.items() if "sample" in k})


def _cast_like_from_pretrained(model, dtype, keep_fp32):
    """``from_pretrained(torch_dtype=dtype)`` + ``_keep_in_fp32_modules``: every floating parameter/buffer is cast to
    ``dtype`` except those whose name contains one of the keep patterns (transformer_wan.py:393)."""
    for name, p in model.named_parameters():
        p.data = p.data.to(torch.float32 if any(k in name for k in keep_fp32) else dtype)
    return model


def bf16_golden():
    """The reference classes run IN BF16 on CPU (weights cast as from_pretrained(torch_dtype=bf16) does, keeping
    ``_keep_in_fp32_modules`` in fp32), so that every bf16 cast point of SURVEY.md §9 is live. The oracle in bf16 mode
    must reproduce these (tests/test_oracle_golden.py); fixtures hold the bf16 samples + per-block taps."""
    os.chdir(REF)
    from architecture.cogvideox_transformer_3d import CogVideoXTransformer3DModel
    from architecture.embeddings import get_3d_rotary_pos_embed
    from architecture.transformer_wan import WanTransformer3DModel

    out = {}
    for name, cfg, shape in (("tiny", synth.WAN_TINY, (5, 16, 16)), ("small", synth.WAN_SMALL, (3, 16, 16))):
        model = WanTransformer3DModel(**cfg).eval()
        model.load_state_dict(synth.make_wan_state_dict(cfg, seed=0), strict=True)
        keep = list(WanTransformer3DModel._keep_in_fp32_modules)
        assert sorted(keep) == sorted(synth.WAN_KEEP_FP32), keep
        _cast_like_from_pretrained(model, torch.bfloat16, keep)
        taps = {}
        hooks = []
        for i, blk in enumerate(model.blocks):
            hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.out", o.clone())))
            # attention outputs / conditioning rows: every 8th token row (ROW_STEP) keeps the fixture small
            hooks.append(blk.attn1.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.attn1", o[:, ::8].clone())))
            hooks.append(blk.attn2.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"blocks.{i}.attn2", o[:, ::8].clone())))
        hooks.append(model.condition_embedder.register_forward_hook(
            lambda m, a, o: taps.update(temb=(o[0][:, ::8] if o[0].dim() == 3 else o[0]).clone(),
                                        timestep_proj=(o[1][:, ::8] if o[1].dim() == 3 else o[1]).clone(),
                                        text=o[2].clone())))
        f, h, w = shape
        for mode in ("per_token", "scalar"):
            hidden, ts, text = synth.make_wan_inputs(cfg, f, h, w, n_id=1, text_len=16, text_true_len=11,
                                                     per_token_timestep=(mode == "per_token"), dtype=torch.bfloat16)
            with torch.no_grad():
                y = model(hidden_states=hidden, timestep=ts, encoder_hidden_states=text, return_dict=False)[0]
            assert y.dtype == torch.bfloat16
            out[f"wan.{name}.{mode}.sample"] = y.clone()
            for k, v in taps.items():
                out[f"wan.{name}.{mode}.{k}"] = v.clone()
        for hk in hooks:
            hk.remove()

    cfg = synth.COG_TINY
    model = CogVideoXTransformer3DModel(**cfg).eval()
    model.load_state_dict(synth.make_cog_state_dict(cfg, seed=0), strict=True)
    model = model.to(torch.bfloat16)  # CogVideoX keeps nothing in fp32
    lat_f = (cfg["sample_frames"] - 1) // cfg["temporal_compression_ratio"] + 1
    h, w = cfg["sample_height"], cfg["sample_width"]
    hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=2, dtype=torch.bfloat16)
    gh, gw = h // cfg["patch_size"], w // cfg["patch_size"]
    cos, sin = get_3d_rotary_pos_embed(cfg["attention_head_dim"], ((0, 0), (gh, gw)), (gh, gw), lat_f)
    cos = torch.cat([cos, cos[: gh * gw]], dim=0)
    sin = torch.cat([sin, sin[: gh * gw]], dim=0)
    taps = {}
    hooks = [blk.register_forward_hook(lambda m, a, o, i=i: taps.update({f"blocks.{i}.out": o[0].clone(),
                                                                         f"blocks.{i}.enc": o[1].clone()}))
             for i, blk in enumerate(model.transformer_blocks)]
    with torch.no_grad():
        y = model(hidden_states=hidden, encoder_hidden_states=text, timestep=ts, image_rotary_emb=(cos, sin),
                  return_dict=False)[0]
    assert y.dtype == torch.bfloat16
    out["cog.tiny.sample"] = y.clone()
    for k, v in taps.items():
        out[f"cog.tiny.{k}"] = v.clone()
    for hk in hooks:
        hk.remove()
    torch.save(out, os.path.join(HERE, "bf16_golden.pt"))
    print("wrote bf16_golden.pt:", {k: (tuple(v.shape), str(v.dtype)) for k, v in out.items() if k.endswith("sample")})


def vae_golden():
    """The reference AutoencoderKLWan (Wan2.2 form: is_residual, patch_size 2) executed on CPU in fp32 — its own chunked
    encode / decode with feat_cache — on the small VAE_TINY config: decode of a 3-latent-frame latent (9 frames out),
    encode of a 9-frame clip, and decode of a single latent frame (the first_chunk-only path)."""
    os.chdir(REF)
    from architecture.autoencoder_kl_wan import AutoencoderKLWan

    cfg = dict(synth.VAE_TINY)
    torch.manual_seed(0)
    vae = AutoencoderKLWan(**cfg, latents_mean=[0.0] * cfg["z_dim"], latents_std=[1.0] * cfg["z_dim"]).eval()
    sd = synth.make_vae_state_dict(cfg, seed=0)
    assert set(sd) == set(vae.state_dict()), sorted(set(sd) ^ set(vae.state_dict()))[:10]
    vae.load_state_dict(sd, strict=True)
    z, x = synth.make_vae_inputs(cfg, 3, 4, 6, seed=5)
    out = {}
    with torch.no_grad():
        out["decode.sample"] = vae.decode(z, return_dict=False)[0].clone()
        out["decode1.sample"] = vae.decode(z[:, :, :1], return_dict=False)[0].clone()
        post = vae.encode(x).latent_dist
        out["encode.parameters"] = post.parameters.clone()
        assert torch.equal(post.mode(), post.parameters[:, : cfg["z_dim"]])
    torch.save(out, os.path.join(HERE, "vae_golden.pt"))
    print("wrote vae_golden.pt:", {k: tuple(v.shape) for k, v in out.items()})


def pipeline_golden():
    """The reference pipeline file itself — ``WanImageToVideoPipeline.prepare_latents`` and ``__call__``
    (pipelines/pipeline_wan_i2v_motion_FrameINO.py, Wan2.2 ``expand_timesteps`` branch) — executed on CPU in fp32 around
    the reference transformer and the reference VAE (tiny configs), with pre-computed prompt embeddings and the shim's
    flow-match Euler scheduler (shift 5). Pixel-space inputs come from synth.make_pipeline_inputs, so the fixture holds
    only outputs: the five prepare_latents tensors, the final latents and the decoded video."""
    os.chdir(REF)
    from architecture.autoencoder_kl_wan import AutoencoderKLWan
    from architecture.transformer_wan import WanTransformer3DModel
    from diffusers.schedulers import FlowMatchEulerDiscreteScheduler
    from pipelines.pipeline_wan_i2v_motion_FrameINO import WanImageToVideoPipeline

    vcfg = synth.with_latent_stats(synth.VAE_TINY)
    torch.manual_seed(0)
    vae = AutoencoderKLWan(**vcfg).eval()
    vae.load_state_dict(synth.make_vae_state_dict(synth.VAE_TINY, seed=1), strict=True)
    tf = WanTransformer3DModel(**synth.WAN_TINY).eval()
    tf.load_state_dict(synth.make_wan_state_dict(synth.WAN_TINY, seed=0), strict=True)
    pipe = WanImageToVideoPipeline(tokenizer=None, text_encoder=None, vae=vae,
                                   scheduler=FlowMatchEulerDiscreteScheduler(shift=5.0), transformer=tf,
                                   expand_timesteps=True)
    h, w, f = 64, 96, 9
    inp = synth.make_pipeline_inputs(vcfg, 64, num_frames=f, height=h, width=w, n_id=1)
    out = {}
    with torch.no_grad():
        image = pipe.video_processor.preprocess(inp["image"], height=h, width=w).to(torch.float32)
        names = ["latents", "latent_condition", "traj_latents", "ID_latent_condition", "first_frame_mask"]
        got = pipe.prepare_latents(image, inp["traj_tensor"], inp["ID_tensor"], 1, vcfg["z_dim"], h, w, f,
                                   torch.float32, torch.device("cpu"), None, inp["latents"].clone(), None)
        for n, t in zip(names, got):
            out["prepare." + n] = t.clone()
        common = dict(image=inp["image"], traj_tensor=inp["traj_tensor"], prompt_embeds=inp["prompt_embeds"],
                      negative_prompt_embeds=inp["negative_prompt_embeds"], height=h, width=w, num_frames=f,
                      num_inference_steps=4, guidance_scale=5.0)
        out["call.latents"] = pipe(ID_tensor=inp["ID_tensor"].clone(), latents=inp["latents"].clone(),
                                   output_type="latent", **common).frames.clone()
        out["call.video"] = pipe(ID_tensor=inp["ID_tensor"].clone(), latents=inp["latents"].clone(), output_type="pt",
                                 **common).frames.clone()
        # no ID frame (the `ID_tensor.shape[2] == 0` branch, :489/:520) and no guidance (a single forward per step)
        empty = inp["ID_tensor"][:, :, :0]
        out["call_noid_nocfg.latents"] = pipe(ID_tensor=empty, latents=inp["latents"].clone(), output_type="latent",
                                              **{**common, "guidance_scale": 1.0}).frames.clone()
    torch.save(out, os.path.join(HERE, "pipeline_golden.pt"))
    print("wrote pipeline_golden.pt:", {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["wan", "cog", "bf16", "vae", "pipeline"]
    if "pipeline" in which:
        pipeline_golden()
    if "vae" in which:
        vae_golden()
    if "bf16" in which:
        bf16_golden()
    if "wan" in which:
        wan_golden()
    if "cog" in which:
        cog_golden()
