"""Seeded synthetic weights and inputs for the configurations BASELINE.json names (no checkpoints or datasets are
reachable offline). Pure functions of (config, seed): the same recipe feeds the native model, the CPU oracle and the
golden-vector generator, so all three see identical tensors.

State-dict keys follow the diffusers layout the reference checkpoints use.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

# ---- Wan -------------------------------------------------------------------------------------------------------
WAN_TINY = dict(patch_size=(1, 2, 2), num_attention_heads=8, attention_head_dim=32, in_channels=32, out_channels=16,
                text_dim=64, freq_dim=256, ffn_dim=1024, num_layers=2, cross_attn_norm=True,
                qk_norm="rms_norm_across_heads", eps=1e-6, rope_max_seq_len=1024)
# head_dim 128 at a size the CPU oracle still finishes in seconds (used by the GPU parity tests)
WAN_SMALL = dict(patch_size=(1, 2, 2), num_attention_heads=4, attention_head_dim=128, in_channels=32, out_channels=16,
                 text_dim=128, freq_dim=256, ffn_dim=2048, num_layers=2, cross_attn_norm=True,
                 qk_norm="rms_norm_across_heads", eps=1e-6, rope_max_seq_len=1024)
WAN22_5B = dict(patch_size=(1, 2, 2), num_attention_heads=24, attention_head_dim=128, in_channels=96, out_channels=48,
                text_dim=4096, freq_dim=256, ffn_dim=14336, num_layers=30, cross_attn_norm=True,
                qk_norm="rms_norm_across_heads", eps=1e-6, rope_max_seq_len=1024)

WAN_KEEP_FP32 = ("time_embedder", "scale_shift_table", "norm1", "norm2", "norm3")  # transformer_wan.py:393


def wan_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    pt, ph, pw = cfg["patch_size"]
    f = cfg["ffn_dim"]
    shapes: Dict[str, Tuple[int, ...]] = {
        "patch_embedding.weight": (d, cfg["in_channels"], pt, ph, pw),
        "patch_embedding.bias": (d,),
        "condition_embedder.time_embedder.linear_1.weight": (d, cfg["freq_dim"]),
        "condition_embedder.time_embedder.linear_1.bias": (d,),
        "condition_embedder.time_embedder.linear_2.weight": (d, d),
        "condition_embedder.time_embedder.linear_2.bias": (d,),
        "condition_embedder.time_proj.weight": (6 * d, d),
        "condition_embedder.time_proj.bias": (6 * d,),
        "condition_embedder.text_embedder.linear_1.weight": (d, cfg["text_dim"]),
        "condition_embedder.text_embedder.linear_1.bias": (d,),
        "condition_embedder.text_embedder.linear_2.weight": (d, d),
        "condition_embedder.text_embedder.linear_2.bias": (d,),
        "scale_shift_table": (1, 2, d),
        "proj_out.weight": (cfg["out_channels"] * pt * ph * pw, d),
        "proj_out.bias": (cfg["out_channels"] * pt * ph * pw,),
    }
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}"
        shapes[f"{p}.scale_shift_table"] = (1, 6, d)
        for a in ("attn1", "attn2"):
            for proj in ("to_q", "to_k", "to_v", "to_out.0"):
                shapes[f"{p}.{a}.{proj}.weight"] = (d, d)
                shapes[f"{p}.{a}.{proj}.bias"] = (d,)
            shapes[f"{p}.{a}.norm_q.weight"] = (d,)
            shapes[f"{p}.{a}.norm_k.weight"] = (d,)
        if cfg.get("cross_attn_norm", True):
            shapes[f"{p}.norm2.weight"] = (d,)
            shapes[f"{p}.norm2.bias"] = (d,)
        shapes[f"{p}.ffn.net.0.proj.weight"] = (f, d)
        shapes[f"{p}.ffn.net.0.proj.bias"] = (f,)
        shapes[f"{p}.ffn.net.2.weight"] = (d, f)
        shapes[f"{p}.ffn.net.2.bias"] = (d,)
    return shapes


def _fill(name: str, shape: Tuple[int, ...], gen: torch.Generator, device) -> torch.Tensor:
    """Value recipe: GEMM weights N(0, 1/fan_in), biases N(0, 0.02^2), norm gains 1 + 0.1 N(0,1),
    scale_shift_table N(0,1)/sqrt(D) (as transformer_wan.py:306,450), positional tables N(0, 0.02^2)."""
    if name.endswith("scale_shift_table"):
        return torch.randn(shape, generator=gen, device=device) / math.sqrt(shape[-1])
    if "pos_embedding" in name:
        return torch.randn(shape, generator=gen, device=device) * 0.02
    if name.endswith(".bias"):
        return torch.randn(shape, generator=gen, device=device) * 0.02
    if len(shape) == 1 or name.endswith(".gamma"):  # norm gains
        return 1.0 + 0.1 * torch.randn(shape, generator=gen, device=device)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=gen, device=device) / math.sqrt(fan_in)


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, dtype: torch.dtype = torch.float32,
                    keep_fp32: Tuple[str, ...] = (), device="cpu") -> Dict[str, torch.Tensor]:
    """Fills parameters in sorted-key order from one seeded generator. ``dtype`` applies to everything except keys
    containing one of ``keep_fp32`` (diffusers' ``_keep_in_fp32_modules``)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        t = _fill(name, shapes[name], gen, device)
        keep = any(k in name for k in keep_fp32)
        sd[name] = t if (keep or dtype == torch.float32) else t.to(dtype)
    return sd


def make_wan_state_dict(cfg: dict, seed: int = 0, dtype: torch.dtype = torch.float32, device="cpu"):
    return make_state_dict(wan_param_shapes(cfg), seed, dtype, WAN_KEEP_FP32 if dtype != torch.float32 else (), device)


def make_wan_inputs(cfg: dict, latent_frames: int, height: int, width: int, n_id: int = 1, text_len: int = 16,
                    text_true_len: Optional[int] = None, t_value: float = 500.0, batch: int = 1, seed: int = 0,
                    per_token_timestep: bool = True, dtype: torch.dtype = torch.float32, device="cpu"):
    """Inputs shaped like the FrameINO Wan sampler builds them (pipelines/pipeline_wan_i2v_motion_FrameINO.py:826-858):
    ``[noisy latents ‖ ID frames]`` frame-wise, trajectory latents channel-wise with zeros on the ID frames; text rows
    past the true prompt length are exact zeros (:235-238); per-token timesteps are 0 on latent frame 0 and ``t``
    elsewhere including the ID tokens (:832-843)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1000)
    c_half = cfg["in_channels"] // 2
    f_all = latent_frames + n_id
    lat = torch.randn(batch, c_half, f_all, height, width, generator=gen, device=device)
    traj = torch.randn(batch, c_half, f_all, height, width, generator=gen, device=device)
    if n_id > 0:
        traj[:, :, latent_frames:] = 0
    hidden = torch.cat([lat, traj], dim=1).to(dtype)
    text = torch.randn(batch, text_len, cfg["text_dim"], generator=gen, device=device)
    if text_true_len is not None:
        text[:, text_true_len:] = 0
    text = text.to(dtype)
    pt, ph, pw = cfg["patch_size"]
    tokens_per_frame = (height // ph) * (width // pw)
    n_tokens = (f_all // pt) * tokens_per_frame
    if per_token_timestep:
        ts = torch.full((batch, n_tokens), float(t_value), device=device)
        ts[:, :tokens_per_frame] = 0.0
    else:
        ts = torch.full((batch,), float(t_value), device=device)
    return hidden, ts, text


# ---- CogVideoX -------------------------------------------------------------------------------------------------
COG_TINY = dict(num_attention_heads=4, attention_head_dim=64, in_channels=48, out_channels=16, time_embed_dim=64,
                text_embed_dim=64, num_layers=2, sample_width=16, sample_height=12, sample_frames=9, patch_size=2,
                temporal_compression_ratio=4, max_text_seq_length=10, norm_eps=1e-5,
                use_rotary_positional_embeddings=True, use_learned_positional_embeddings=True, use_FrameIn=True)
COG_5B_I2V = dict(num_attention_heads=48, attention_head_dim=64, in_channels=48, out_channels=16, time_embed_dim=512,
                  text_embed_dim=4096, num_layers=42, sample_width=90, sample_height=60, sample_frames=49,
                  patch_size=2, temporal_compression_ratio=4, max_text_seq_length=226, norm_eps=1e-5,
                  use_rotary_positional_embeddings=True, use_learned_positional_embeddings=True, use_FrameIn=True)


def cog_param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    hd = cfg["attention_head_dim"]
    p = cfg["patch_size"]
    te = cfg["time_embed_dim"]
    frames = (cfg["sample_frames"] - 1) // cfg["temporal_compression_ratio"] + 1
    n_patches = (cfg["sample_height"] // p) * (cfg["sample_width"] // p) * frames
    shapes: Dict[str, Tuple[int, ...]] = {
        "patch_embed.proj.weight": (d, cfg["in_channels"], p, p),
        "patch_embed.proj.bias": (d,),
        "patch_embed.text_proj.weight": (d, cfg["text_embed_dim"]),
        "patch_embed.text_proj.bias": (d,),
        "time_embedding.linear_1.weight": (te, d),
        "time_embedding.linear_1.bias": (te,),
        "time_embedding.linear_2.weight": (te, te),
        "time_embedding.linear_2.bias": (te,),
        "norm_final.weight": (d,),
        "norm_final.bias": (d,),
        "norm_out.linear.weight": (2 * d, te),
        "norm_out.linear.bias": (2 * d,),
        "norm_out.norm.weight": (d,),
        "norm_out.norm.bias": (d,),
        "proj_out.weight": (p * p * cfg["out_channels"], d),
        "proj_out.bias": (p * p * cfg["out_channels"],),
    }
    if cfg.get("use_learned_positional_embeddings", False):
        shapes["patch_embed.pos_embedding"] = (1, cfg["max_text_seq_length"] + n_patches, d)
    for i in range(cfg["num_layers"]):
        b = f"transformer_blocks.{i}"
        for nrm in ("norm1", "norm2"):
            shapes[f"{b}.{nrm}.linear.weight"] = (6 * d, te)
            shapes[f"{b}.{nrm}.linear.bias"] = (6 * d,)
            shapes[f"{b}.{nrm}.norm.weight"] = (d,)
            shapes[f"{b}.{nrm}.norm.bias"] = (d,)
        for proj in ("to_q", "to_k", "to_v", "to_out.0"):
            shapes[f"{b}.attn1.{proj}.weight"] = (d, d)
            shapes[f"{b}.attn1.{proj}.bias"] = (d,)
        for nrm in ("norm_q", "norm_k"):
            shapes[f"{b}.attn1.{nrm}.weight"] = (hd,)
            shapes[f"{b}.attn1.{nrm}.bias"] = (hd,)
        shapes[f"{b}.ff.net.0.proj.weight"] = (4 * d, d)
        shapes[f"{b}.ff.net.0.proj.bias"] = (4 * d,)
        shapes[f"{b}.ff.net.2.weight"] = (d, 4 * d)
        shapes[f"{b}.ff.net.2.bias"] = (d,)
    return shapes


def make_cog_state_dict(cfg: dict, seed: int = 0, dtype: torch.dtype = torch.float32, device="cpu"):
    sd = make_state_dict(cog_param_shapes(cfg), seed, dtype, (), device)
    if "patch_embed.pos_embedding" in sd:
        sd["patch_embed.pos_embedding"][:, : cfg["max_text_seq_length"]] = 0  # text rows are zeros, embeddings.py:710-713
    return sd


def make_cog_inputs(cfg: dict, latent_frames: int, height: int, width: int, n_id: int = 1, batch: int = 2,
                    text_len: Optional[int] = None, t_value: float = 500.0, seed: int = 0,
                    dtype: torch.dtype = torch.float32, device="cpu"):
    """[B, F+n_id, 48, H, W] = cat(noisy‖ID, image‖0, traj‖0) on the channel axis
    (pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:856-881); scalar timestep per sample."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 2000)
    c3 = cfg["in_channels"] // 3
    f_all = latent_frames + n_id
    parts = [torch.randn(batch, f_all, c3, height, width, generator=gen, device=device) for _ in range(3)]
    if n_id > 0:
        parts[1][:, latent_frames:] = 0
        parts[2][:, latent_frames:] = 0
    hidden = torch.cat(parts, dim=2).to(dtype)
    tl = cfg["max_text_seq_length"] if text_len is None else text_len
    text = torch.randn(batch, tl, cfg["text_embed_dim"], generator=gen, device=device).to(dtype)
    ts = torch.full((batch,), float(t_value), device=device)
    return hidden, ts, text


# ---- full-size construction without a CPU copy ------------------------------------------------------------------
def build_wan_on_device(cfg: dict, seed: int = 0, device="cuda", dtype: torch.dtype = torch.bfloat16):
    """Random-init ``frameino_b200.WanTransformer3DModel`` materialised directly on ``device`` (the 5B model would
    need 20 GB of host RAM otherwise): parameters are created on the meta device, then filled one by one from the
    seeded recipe with the ``from_pretrained(torch_dtype=bf16)`` + ``_keep_in_fp32_modules`` dtype policy."""
    from .wan import WanRotaryPosEmbed, WanTransformer3DModel

    with torch.device("meta"):
        model = WanTransformer3DModel(**cfg)
    shapes = wan_param_shapes(cfg)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    model.to_empty(device=device)
    params = dict(model.named_parameters())
    assert set(params) == set(shapes), "state-dict layout drifted from synth.wan_param_shapes"
    with torch.no_grad():
        for name in sorted(shapes):
            keep = any(k in name for k in WAN_KEEP_FP32)
            t = _fill(name, shapes[name], gen, device)
            p = params[name]
            p.data = t.to(torch.float32 if keep else dtype)
    model.rope = WanRotaryPosEmbed(cfg["attention_head_dim"], cfg["patch_size"], cfg["rope_max_seq_len"]).to(device)
    return model.eval()


def build_vae_on_device(cfg: dict, seed: int = 0, device="cuda"):
    """Random-init ``frameino_b200.vae.AutoencoderKLWan`` (fp32 parameters, as the reference loads its VAE: app.py:157)
    materialised directly on ``device`` with its bf16 weight packs prepared."""
    from .vae import AutoencoderKLWan

    with torch.device("meta"):
        m = AutoencoderKLWan(**cfg)
    m.to_empty(device=device)
    shapes = vae_param_shapes(cfg)
    params = dict(m.named_parameters())
    assert set(params) == set(shapes), "state-dict layout drifted from synth.vae_param_shapes"
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name in sorted(shapes):
            params[name].data = _fill(name, shapes[name], gen, device)
    return m.eval().prepare()


def build_cog_on_device(cfg: dict, seed: int = 0, device="cuda", dtype: torch.dtype = torch.bfloat16):
    """Random-init ``frameino_b200.CogVideoXTransformer3DModel`` materialised directly on ``device``."""
    from .cogvideox import CogVideoXTransformer3DModel

    with torch.device("meta"):
        model = CogVideoXTransformer3DModel(**cfg)
    shapes = cog_param_shapes(cfg)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    model.to_empty(device=device)
    tensors = dict(model.named_parameters())
    tensors.update(dict(model.named_buffers()))
    assert set(shapes) <= set(tensors), "state-dict layout drifted from synth.cog_param_shapes"
    with torch.no_grad():
        for name in sorted(shapes):
            t = _fill(name, shapes[name], gen, device).to(dtype)
            if name == "patch_embed.pos_embedding":
                t[:, : cfg["max_text_seq_length"]] = 0
            tensors[name].data = t
    return model.eval()


def cog_rope_tables(head_dim: int, grid_h: int, grid_w: int, frames: int, n_id: int = 1, device="cpu"):
    """(cos, sin) [frames*gh*gw + n_id*gh*gw, head_dim] fp32 as the FrameINO CogVideoX pipeline builds them
    (pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:540-584, :834-839; embeddings.py:864-962): linspace grids,
    bands t d/4 | h 3d/8 | w 3d/8, interleaved-repeated, ID frames = copies of frame 0."""
    def band(dim, pos):
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
        ang = torch.outer(pos, freqs)
        return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)

    gh = torch.linspace(0, grid_h * (grid_h - 1) / grid_h, grid_h, dtype=torch.float32)
    gw = torch.linspace(0, grid_w * (grid_w - 1) / grid_w, grid_w, dtype=torch.float32)
    gt = torch.linspace(0, frames * (frames - 1) / frames, frames, dtype=torch.float32)
    t, h, w = band(head_dim // 4, gt), band(head_dim // 8 * 3, gh), band(head_dim // 8 * 3, gw)
    out = []
    for i in (0, 1):
        tab = torch.cat([t[i][:, None, None, :].expand(-1, grid_h, grid_w, -1),
                         h[i][None, :, None, :].expand(frames, -1, grid_w, -1),
                         w[i][None, None, :, :].expand(frames, grid_h, -1, -1)], dim=-1).reshape(frames * grid_h * grid_w, -1)
        if n_id:
            tab = torch.cat([tab] + [tab[: grid_h * grid_w]] * n_id, dim=0)
        out.append(tab.contiguous().to(device))
    return out[0], out[1]


# ---- Wan VAE -------------------------------------------------------------------------------
# Wan2.2-TI2V-5B VAE (upstream HF config, recalled — SURVEY.md §10) and a small config of the same shape for tests
WAN22_VAE = dict(base_dim=160, decoder_base_dim=256, z_dim=48, dim_mult=[1, 2, 4, 4], num_res_blocks=2,
                 temperal_downsample=[False, True, True], is_residual=True, in_channels=12, out_channels=12, patch_size=2,
                 scale_factor_temporal=4, scale_factor_spatial=16)
VAE_TINY = dict(base_dim=32, decoder_base_dim=64, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2,
                temperal_downsample=[False, True, True], is_residual=True, in_channels=12, out_channels=12, patch_size=2,
                scale_factor_temporal=4, scale_factor_spatial=16)


def _vae_dims(cfg: dict, decoder: bool):
    mult = list(cfg["dim_mult"])
    if decoder:
        dim = cfg.get("decoder_base_dim") or cfg["base_dim"]
        return [dim * u for u in [mult[-1]] + mult[::-1]]  # autoencoder_kl_wan.py:821
    return [cfg["base_dim"] * u for u in [1] + mult]  # :545


def vae_param_shapes(cfg: dict) -> Dict[str, tuple]:
    """State-dict layout of the reference AutoencoderKLWan (is_residual) — names from :505-584, :783-872, :1022-1050."""
    shapes: Dict[str, tuple] = {}
    z = cfg["z_dim"]

    def conv3(name, cin, cout, k):
        shapes[name + ".weight"] = (cout, cin, *k)
        shapes[name + ".bias"] = (cout,)

    def res(name, cin, cout):
        shapes[name + ".norm1.gamma"] = (cin, 1, 1, 1)
        conv3(name + ".conv1", cin, cout, (3, 3, 3))
        shapes[name + ".norm2.gamma"] = (cout, 1, 1, 1)
        conv3(name + ".conv2", cout, cout, (3, 3, 3))
        if cin != cout:
            conv3(name + ".conv_shortcut", cin, cout, (1, 1, 1))

    def mid(name, c):
        res(name + ".resnets.0", c, c)
        shapes[name + ".attentions.0.norm.gamma"] = (c, 1, 1)
        shapes[name + ".attentions.0.to_qkv.weight"] = (3 * c, c, 1, 1)
        shapes[name + ".attentions.0.to_qkv.bias"] = (3 * c,)
        shapes[name + ".attentions.0.proj.weight"] = (c, c, 1, 1)
        shapes[name + ".attentions.0.proj.bias"] = (c,)
        res(name + ".resnets.1", c, c)

    n = len(cfg["dim_mult"])
    # encoder
    dims = _vae_dims(cfg, False)
    conv3("encoder.conv_in", cfg["in_channels"], dims[0], (3, 3, 3))
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        p = f"encoder.down_blocks.{i}"
        c = cin
        for j in range(cfg["num_res_blocks"]):
            res(f"{p}.resnets.{j}", c, cout)
            c = cout
        if i != n - 1:
            shapes[f"{p}.downsampler.resample.1.weight"] = (cout, cout, 3, 3)
            shapes[f"{p}.downsampler.resample.1.bias"] = (cout,)
            if cfg["temperal_downsample"][i]:
                conv3(f"{p}.downsampler.time_conv", cout, cout, (3, 1, 1))
    mid("encoder.mid_block", dims[-1])
    shapes["encoder.norm_out.gamma"] = (dims[-1], 1, 1, 1)
    conv3("encoder.conv_out", dims[-1], 2 * z, (3, 3, 3))
    conv3("quant_conv", 2 * z, 2 * z, (1, 1, 1))
    conv3("post_quant_conv", z, z, (1, 1, 1))
    # decoder
    dims = _vae_dims(cfg, True)
    t_up = list(cfg["temperal_downsample"])[::-1]
    conv3("decoder.conv_in", z, dims[0], (3, 3, 3))
    mid("decoder.mid_block", dims[0])
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        p = f"decoder.up_blocks.{i}"
        c = cin
        for j in range(cfg["num_res_blocks"] + 1):
            res(f"{p}.resnets.{j}", c, cout)
            c = cout
        if i != n - 1:
            shapes[f"{p}.upsampler.resample.1.weight"] = (cout, cout, 3, 3)
            shapes[f"{p}.upsampler.resample.1.bias"] = (cout,)
            if t_up[i]:
                conv3(f"{p}.upsampler.time_conv", cout, 2 * cout, (3, 1, 1))
    shapes["decoder.norm_out.gamma"] = (dims[-1], 1, 1, 1)
    conv3("decoder.conv_out", dims[-1], cfg["out_channels"], (3, 3, 3))
    return shapes


def make_vae_state_dict(cfg: dict, seed: int = 0, dtype: torch.dtype = torch.float32, device="cpu"):
    return make_state_dict(vae_param_shapes(cfg), seed, dtype, (), device)


def make_vae_inputs(cfg: dict, latent_frames: int = 3, h: int = 4, w: int = 6, seed: int = 5, device="cpu"):
    """(z, x): a latent [1, z_dim, T, h, w] for decode and a clip [1, 3, 1 + 4 (T - 1), h*s, w*s] in [-1, 1] for encode
    (s = scale_factor_spatial), from one seeded generator."""
    g = torch.Generator(device=device).manual_seed(seed)
    s_ = cfg["scale_factor_spatial"]
    z = torch.randn(1, cfg["z_dim"], latent_frames, h, w, generator=g, device=device)
    x = torch.randn(1, 3, 1 + 4 * (latent_frames - 1), h * s_, w * s_, generator=g, device=device).clamp(-1, 1)
    return z, x


def with_latent_stats(cfg: dict, seed: int = 11) -> dict:
    """A copy of a VAE config with seeded ``latents_mean`` / ``latents_std`` lists (the released values are in the
    HF config, not under /root/reference; the pipeline only needs them to be per-channel numbers)."""
    g = torch.Generator().manual_seed(seed)
    z = cfg["z_dim"]
    out = dict(cfg)
    out["latents_mean"] = (0.3 * torch.randn(z, generator=g)).tolist()
    out["latents_std"] = (0.5 + torch.rand(z, generator=g)).tolist()
    return out


def make_pipeline_inputs(vae_cfg: dict, text_dim: int, num_frames: int, height: int, width: int, n_id: int = 1,
                         text_len: int = 16, seed: int = 9, batch: int = 1):
    """Pixel-space inputs of the Wan FrameINO pipeline call (app.py:705-714): first-frame canvas [1, 3, H, W] in
    [-1, 1], trajectory video [F, 3, H, W], ID image(s) [1, 3, n_id, H, W], prompt / negative prompt embeddings."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(1, 3, height, width, generator=g) * 2 - 1
    image[:, :, : height // 4] = -1.0  # an outpainting border, as the FrameINO canvas has
    traj = (torch.rand(num_frames, 3, height, width, generator=g) < 0.05).float() * 2 - 1
    ident = torch.rand(1, 3, n_id, height, width, generator=g) * 2 - 1 if n_id else None
    pos = torch.randn(batch, text_len, text_dim, generator=g)
    pos[:, text_len * 3 // 4:] = 0  # zero rows after the true length (pipeline :235-238)
    neg = torch.zeros(batch, text_len, text_dim)
    neg[:, :2] = torch.randn(batch, 2, text_dim, generator=g)
    lat_shape = (batch, vae_cfg["z_dim"], (num_frames - 1) // vae_cfg["scale_factor_temporal"] + 1,
                 height // vae_cfg["scale_factor_spatial"], width // vae_cfg["scale_factor_spatial"])
    latents = torch.randn(lat_shape, generator=g)
    return dict(image=image, traj_tensor=traj, ID_tensor=ident, prompt_embeds=pos, negative_prompt_embeds=neg,
                latents=latents)
