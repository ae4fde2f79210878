"""CPU oracle (test infrastructure) — CogVideoX-5B-I2V FrameINO transformer forward, restated from
``/root/reference/architecture/cogvideox_transformer_3d.py``, ``attention_processor.py:2805-2877`` and
``embeddings.py`` as pure functions over a diffusers-layout state dict. See oracle/__init__.py for the pinning status.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .wan_oracle import linear, rope_tables_1d, sdpa, sinusoidal_embedding


def cog_rope_3d(head_dim: int, grid_hw: Tuple[int, int], temporal_size: int, n_id_frames: int = 0):
    """get_3d_rotary_pos_embed (embeddings.py:864-962) with crops ((0,0),(gh,gw)) and grid_type 'linspace' as the
    1.0-checkpoint branch of the pipeline calls it (pipeline_cogvideox_i2v_motion_FrameINO.py:540-584), then the ID
    frame rows = copy of frame-0 rows (:834-839). Returns (cos, sin) each [(T+n_id)*gh*gw, head_dim] fp32."""
    gh, gw = grid_hw
    grid_h = torch.linspace(0, gh * (gh - 1) / gh, gh, dtype=torch.float32)  # :898-900
    grid_w = torch.linspace(0, gw * (gw - 1) / gw, gw, dtype=torch.float32)  # :901-903
    grid_t = torch.linspace(0, temporal_size * (temporal_size - 1) / temporal_size, temporal_size,
                            dtype=torch.float32)  # :905-907
    dim_t, dim_h, dim_w = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3  # :921-923
    t = rope_tables_1d(dim_t, grid_t, dtype=torch.float32)
    h = rope_tables_1d(dim_h, grid_h, dtype=torch.float32)
    w = rope_tables_1d(dim_w, grid_w, dtype=torch.float32)
    out = []
    for i in (0, 1):
        ft = t[i][:, None, None, :].expand(-1, gh, gw, -1)
        fh = h[i][None, :, None, :].expand(temporal_size, -1, gw, -1)
        fw = w[i][None, None, :, :].expand(temporal_size, gh, -1, -1)
        tab = torch.cat([ft, fh, fw], dim=-1).reshape(temporal_size * gh * gw, -1)
        if n_id_frames:
            tab = torch.cat([tab] + [tab[: gh * gw]] * n_id_frames, dim=0)
        out.append(tab)
    return out[0], out[1]


def apply_cog_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb, use_real=True, unbind_dim=-1 (embeddings.py:1240-1256). x: [B, H, S, d]."""
    cos, sin = cos[None, None], sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def cog_patch_embed(sd, cfg: dict, text: torch.Tensor, video: torch.Tensor) -> torch.Tensor:
    """CogVideoXPatchEmbed.forward (embeddings.py:718-805), patch_size_t=None branch, FrameINO ID-frame append."""
    p = cfg["patch_size"]
    text_e = linear(text, sd, "patch_embed.text_proj")  # :727
    b, f, c, h, w = video.shape
    v = F.conv2d(video.reshape(-1, c, h, w), sd["patch_embed.proj.weight"], sd.get("patch_embed.proj.bias"), stride=p)
    v = v.view(b, f, *v.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)  # :734-738
    emb = torch.cat([text_e, v], dim=1).contiguous()  # :750
    if "patch_embed.pos_embedding" in sd:
        pos = sd["patch_embed.pos_embedding"]
        text_len = text_e.shape[1]
        max_text = cfg["max_text_seq_length"]
        tcr = cfg["temporal_compression_ratio"]
        pre_frames = (f - 1) * tcr + 1
        post_frames = (cfg["sample_frames"] - 1) // tcr + 1
        ph, pw = cfg["sample_height"] // p, cfg["sample_width"] // p
        seq = h * w * f // (p * p)
        if cfg.get("use_FrameIn", False):  # :772-775
            first = (pos.shape[1] - max_text) // (f - 1)
            pos = torch.cat([pos, pos[:, text_len:text_len + first].clone()], dim=1)
        if cfg["sample_height"] != h or cfg["sample_width"] != w or cfg["sample_frames"] != pre_frames:  # :782-798
            if cfg.get("use_FrameIn", False):
                post_frames = post_frames + 1
            d = emb.shape[-1]
            pv = pos[:, text_len:].view(1, post_frames, ph, pw, d).permute(0, 4, 1, 2, 3)
            pv = F.interpolate(pv, size=[post_frames, h // p, w // p], mode="trilinear", align_corners=False)
            pv = pv.permute(0, 2, 3, 4, 1).reshape(1, -1, d)
            pos = torch.cat([pos[:, :text_len], pv], dim=1)[:, : text_len + seq]
        emb = emb + pos.to(emb.dtype)  # :802-803
    return emb


def layer_norm_zero(sd, prefix: str, x, enc, temb, eps: float):
    """diffusers CogVideoXLayerNormZero (upstream)."""
    mods = linear(F.silu(temb), sd, prefix + ".linear").chunk(6, dim=1)
    shift, scale, gate, e_shift, e_scale, e_gate = mods
    w, b = sd.get(prefix + ".norm.weight"), sd.get(prefix + ".norm.bias")
    d = x.shape[-1]
    xn = F.layer_norm(x, (d,), w, b, eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    en = F.layer_norm(enc, (d,), w, b, eps) * (1 + e_scale)[:, None, :] + e_shift[:, None, :]
    return xn, en, gate[:, None, :], e_gate[:, None, :]


def cog_attention(sd, prefix: str, cfg: dict, x, enc, rope, taps=None):
    """CogVideoXAttnProcessor2_0.__call__ (attention_processor.py:2815-2877)."""
    heads, hd = cfg["num_attention_heads"], cfg["attention_head_dim"]
    text_len = enc.size(1)
    s = torch.cat([enc, x], dim=1)  # :2827
    b = s.shape[0]
    q, k, v = (linear(s, sd, f"{prefix}.{n}") for n in ("to_q", "to_k", "to_v"))  # :2837-2839
    q, k, v = (t.view(b, -1, heads, hd).transpose(1, 2) for t in (q, k, v))  # :2844-2846
    q = F.layer_norm(q, (hd,), sd[prefix + ".norm_q.weight"], sd[prefix + ".norm_q.bias"], 1e-6)  # :2848-2851
    k = F.layer_norm(k, (hd,), sd[prefix + ".norm_k.weight"], sd[prefix + ".norm_k.bias"], 1e-6)
    if rope is not None:  # :2855-2860
        q = q.clone()
        k = k.clone()
        q[:, :, text_len:] = apply_cog_rope(q[:, :, text_len:], *rope)
        k[:, :, text_len:] = apply_cog_rope(k[:, :, text_len:], *rope)
    o = sdpa(q, k, v)  # :2863
    o = o.transpose(1, 2).reshape(b, -1, heads * hd)
    o = linear(o, sd, prefix + ".to_out.0")  # :2870
    return o[:, text_len:], o[:, :text_len]  # :2874-2877


def cog_block(sd, i: int, cfg: dict, x, enc, temb, rope, taps=None):
    """CogVideoXBlock.forward (cogvideox_transformer_3d.py:122-161)."""
    p = f"transformer_blocks.{i}"
    eps = cfg.get("norm_eps", 1e-5)
    text_len = enc.size(1)
    xn, en, gate, e_gate = layer_norm_zero(sd, p + ".norm1", x, enc, temb, eps)  # :134
    a_x, a_e = cog_attention(sd, p + ".attn1", cfg, xn, en, rope, taps)  # :139
    x = x + gate * a_x  # :146
    enc = enc + e_gate * a_e  # :147
    if taps is not None:
        taps[f"{p}.after_attn"] = torch.cat([enc, x], dim=1)
    xn, en, gate_ff, e_gate_ff = layer_norm_zero(sd, p + ".norm2", x, enc, temb, eps)  # :150
    h = torch.cat([en, xn], dim=1)  # :155
    h = linear(F.gelu(linear(h, sd, p + ".ff.net.0.proj"), approximate="tanh"), sd, p + ".ff.net.2")  # :156
    x = x + gate_ff * h[:, text_len:]  # :158
    enc = enc + e_gate_ff * h[:, :text_len]  # :159
    if taps is not None:
        taps[f"{p}.out"] = x
        taps[f"{p}.enc"] = enc
    return x, enc


def cog_forward(sd: Dict[str, torch.Tensor], cfg: dict, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                timestep: torch.Tensor, image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                taps: Optional[dict] = None, num_layers: Optional[int] = None) -> torch.Tensor:
    """CogVideoXTransformer3DModel.forward (cogvideox_transformer_3d.py:446-562), ofs branch off (5b-I2V 1.0)."""
    b, f, c, h, w = hidden_states.shape
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    eps = cfg.get("norm_eps", 1e-5)
    t_emb = sinusoidal_embedding(timestep, d, flip_sin_to_cos=True, shift=0.0).to(hidden_states.dtype)  # :478-484
    emb = linear(F.silu(linear(t_emb, sd, "time_embedding.linear_1")), sd, "time_embedding.linear_2")  # :485
    x = cog_patch_embed(sd, cfg, encoder_hidden_states, hidden_states)  # :494
    text_len = encoder_hidden_states.shape[1]
    enc, x = x[:, :text_len], x[:, text_len:]  # :498-500
    n_layers = cfg["num_layers"] if num_layers is None else num_layers
    for i in range(n_layers):  # :503-529
        x, enc = cog_block(sd, i, cfg, x, enc, emb, image_rotary_emb, taps)
    if not cfg.get("use_rotary_positional_embeddings", False):  # :531-538
        x = F.layer_norm(x, (d,), sd.get("norm_final.weight"), sd.get("norm_final.bias"), eps)
    else:
        x = torch.cat([enc, x], dim=1)
        x = F.layer_norm(x, (d,), sd.get("norm_final.weight"), sd.get("norm_final.bias"), eps)[:, text_len:]
    # AdaLayerNorm(chunk_dim=1) (upstream): shift, scale = Linear(SiLU(emb)).chunk(2, dim=1)   (:541)
    shift, scale = linear(F.silu(emb), sd, "norm_out.linear").chunk(2, dim=1)
    x = F.layer_norm(x, (d,), sd.get("norm_out.norm.weight"), sd.get("norm_out.norm.bias"), eps)
    x = x * (1 + scale[:, None, :]) + shift[:, None, :]
    x = linear(x, sd, "proj_out")  # :542
    if taps is not None:
        taps["proj_out"] = x
    p = cfg["patch_size"]
    out = x.reshape(b, f, h // p, w // p, -1, p, p)  # :549
    return out.permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)  # :550
