"""CPU oracle (test infrastructure) — Wan2.2 FrameINO transformer forward, restated from
``/root/reference/architecture/transformer_wan.py`` as pure functions over a diffusers-layout state dict.

Every function cites the reference lines it follows. dtype behaviour follows SURVEY.md §9: the model dtype ``bf`` is the
dtype of the GEMM weights in the state dict (fp32 for the tiny CPU config, bf16 for the GPU configs); fp32 islands are
reproduced where the reference has them. Optional ``taps`` dict collects per-layer tensors for parity tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class WanConfig:
    """Mirror of the constructor arguments at transformer_wan.py:397-416."""

    patch_size: Tuple[int, int, int] = (1, 2, 2)
    num_attention_heads: int = 40
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 13824
    num_layers: int = 40
    cross_attn_norm: bool = True
    qk_norm: Optional[str] = "rms_norm_across_heads"
    eps: float = 1e-6
    image_dim: Optional[int] = None
    added_kv_proj_dim: Optional[int] = None
    rope_max_seq_len: int = 1024
    pos_embed_seq_len: Optional[int] = None

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


# the two configurations BASELINE.json names
WAN_TINY = WanConfig(num_attention_heads=8, attention_head_dim=32, in_channels=32, out_channels=16, text_dim=64,
                     freq_dim=256, ffn_dim=1024, num_layers=2)
WAN22_5B = WanConfig(num_attention_heads=24, attention_head_dim=128, in_channels=96, out_channels=48, text_dim=4096,
                     freq_dim=256, ffn_dim=14336, num_layers=30)


# ----------------------------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------------------------
def sinusoidal_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, shift: float = 0.0) -> torch.Tensor:
    """embeddings.py:27-78 (get_timestep_embedding), scale 1, max_period 10000."""
    assert t.dim() == 1
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - shift)
    ang = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def rope_tables_1d(dim: int, positions: torch.Tensor, theta: float = 10000.0, dtype=torch.float64):
    """embeddings.py:1153-1207 with use_real=True, repeat_interleave_real=True."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=dtype)[: dim // 2] / dim))
    ang = torch.outer(positions.to(dtype), freqs)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def wan_rope(cfg: WanConfig, num_frames: int, height: int, width: int):
    """transformer_wan.py:192-253: (cos, sin), each [1, 1, N, head_dim] fp32, token order f-major/h/w."""
    d = cfg.attention_head_dim
    p_t, p_h, p_w = cfg.patch_size
    ppf, pph, ppw = num_frames // p_t, height // p_h, width // p_w
    h_dim = w_dim = 2 * (d // 6)  # :206
    t_dim = d - h_dim - w_dim  # :207
    pos = torch.arange(cfg.rope_max_seq_len)
    tabs = [rope_tables_1d(dim, pos) for dim in (t_dim, h_dim, w_dim)]
    cos_all = torch.cat([t[0] for t in tabs], dim=1)
    sin_all = torch.cat([t[1] for t in tabs], dim=1)
    split = [d - 2 * (d // 3), d // 3, d // 3]  # :233-237 (must agree with the construction)
    outs = []
    for tab in (cos_all, sin_all):
        f, h, w = tab.split(split, dim=1)
        f = f[:ppf].view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
        h = h[:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
        w = w[:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
        outs.append(torch.cat([f, h, w], dim=-1).reshape(1, 1, ppf * pph * ppw, -1))
    return outs[0], outs[1]


def apply_wan_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """transformer_wan.py:75-90. x: [B, H, N, d]."""
    xr = x.view(*x.shape[:-1], -1, 2)
    x1, x2 = xr[..., 0], xr[..., 1]
    c = cos[..., 0::2]
    s = sin[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * c - x2 * s
    out[..., 1::2] = x1 * s + x2 * c
    return out.type_as(x)


def fp32_layer_norm(x: torch.Tensor, weight, bias, eps: float) -> torch.Tensor:
    """diffusers FP32LayerNorm (upstream): layer_norm in fp32, result cast back to the input dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps).to(x.dtype)


def rms_norm(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """diffusers RMSNorm (upstream): fp32 variance, x*rsqrt, cast to a half-precision weight dtype, times weight."""
    in_dtype = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight is not None:
        if weight.dtype in (torch.float16, torch.bfloat16):
            y = y.to(weight.dtype)
        y = y * weight
    else:
        y = y.to(in_dtype)
    return y


def linear(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def sdpa(q, k, v):
    """F.scaled_dot_product_attention(q, k, v), no mask, non causal (transformer_wan.py:108-110)."""
    return F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)


# ----------------------------------------------------------------------------------------------------------------
# attention processor + block
# ----------------------------------------------------------------------------------------------------------------
def wan_attention(sd, prefix: str, cfg: WanConfig, hidden: torch.Tensor, context: Optional[torch.Tensor],
                  rotary: Optional[Tuple[torch.Tensor, torch.Tensor]], taps=None, tap_name: str = "") -> torch.Tensor:
    """WanAttnProcessor2_0.__call__ (transformer_wan.py:43-119) without the dead I2V add_k_proj branch."""
    heads = cfg.num_attention_heads
    ctx = hidden if context is None else context
    q = linear(hidden, sd, prefix + ".to_q")  # :60
    k = linear(ctx, sd, prefix + ".to_k")  # :61
    v = linear(ctx, sd, prefix + ".to_v")  # :62
    if cfg.qk_norm is not None:
        assert cfg.qk_norm == "rms_norm_across_heads"
        q = rms_norm(q, sd[prefix + ".norm_q.weight"], cfg.eps)  # :64-67, attention_processor.py:208-211
        k = rms_norm(k, sd[prefix + ".norm_k.weight"], cfg.eps)
    q = q.unflatten(2, (heads, -1)).transpose(1, 2)  # :69-71
    k = k.unflatten(2, (heads, -1)).transpose(1, 2)
    v = v.unflatten(2, (heads, -1)).transpose(1, 2)
    if rotary is not None:
        q = apply_wan_rope(q, *rotary)  # :89-90
        k = apply_wan_rope(k, *rotary)
    if taps is not None:
        taps[tap_name + ".q"] = q
        taps[tap_name + ".k"] = k
    o = sdpa(q, k, v)  # :108
    o = o.transpose(1, 2).flatten(2, 3).type_as(q)  # :111-112
    if taps is not None:
        taps[tap_name + ".attn"] = o
    return linear(o, sd, prefix + ".to_out.0")  # :117 (to_out[1] is Dropout(0))


def feed_forward(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers FeedForward(activation_fn="gelu-approximate") (upstream): Linear -> GELU(tanh) -> Linear."""
    h = F.gelu(linear(x, sd, prefix + ".net.0.proj"), approximate="tanh")
    return linear(h, sd, prefix + ".net.2")


def wan_block(sd, i: int, cfg: WanConfig, x: torch.Tensor, text: torch.Tensor, temb: torch.Tensor, rotary, taps=None):
    """WanTransformerBlock.forward (transformer_wan.py:308-350)."""
    p = f"blocks.{i}"
    sst = sd[p + ".scale_shift_table"]
    if temb.ndim == 4:  # :315-326  per-token [B, N, 6, D]
        chunks = (sst.unsqueeze(0) + temb.float()).chunk(6, dim=2)
        shift, scale, gate, c_shift, c_scale, c_gate = [c.squeeze(2) for c in chunks]
    else:  # :327-331  [B, 6, D]
        shift, scale, gate, c_shift, c_scale, c_gate = (sst + temb.float()).chunk(6, dim=1)
    eps = cfg.eps
    # 1. self-attention (:334-336)
    h = (fp32_layer_norm(x.float(), None, None, eps) * (1 + scale) + shift).type_as(x)
    if taps is not None:
        taps[f"{p}.norm1"] = h
    a = wan_attention(sd, p + ".attn1", cfg, h, None, rotary, taps, f"{p}.attn1")
    x = (x.float() + a * gate).type_as(x)
    if taps is not None:
        taps[f"{p}.after_attn1"] = x
    # 2. cross-attention (:339-341)
    if cfg.cross_attn_norm:
        h = fp32_layer_norm(x.float(), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps).type_as(x)
    else:
        h = x.float().type_as(x)
    a = wan_attention(sd, p + ".attn2", cfg, h, text, None, taps, f"{p}.attn2")
    x = x + a
    if taps is not None:
        taps[f"{p}.after_attn2"] = x
    # 3. feed-forward (:344-348)
    h = (fp32_layer_norm(x.float(), None, None, eps) * (1 + c_scale) + c_shift).type_as(x)
    f = feed_forward(sd, p + ".ffn", h)
    x = (x.float() + f.float() * c_gate).type_as(x)
    if taps is not None:
        taps[f"{p}.out"] = x
    return x


# ----------------------------------------------------------------------------------------------------------------
# whole forward
# ----------------------------------------------------------------------------------------------------------------
def wan_condition_embedder(sd, cfg: WanConfig, timestep: torch.Tensor, text: torch.Tensor, ts_seq_len: Optional[int]):
    """WanTimeTextImageEmbedding.forward (transformer_wan.py:168-189), image branch unused (image_dim None)."""
    p = "condition_embedder"
    t = sinusoidal_embedding(timestep, cfg.freq_dim, flip_sin_to_cos=True, shift=0.0)  # :175
    if ts_seq_len is not None:
        t = t.unflatten(0, (1, ts_seq_len))  # :177 (B = 1 only)
    te_dtype = sd[p + ".time_embedder.linear_1.weight"].dtype
    t = t.to(te_dtype)  # :179-181
    temb = linear(F.silu(linear(t, sd, p + ".time_embedder.linear_1")), sd, p + ".time_embedder.linear_2")
    temb = temb.type_as(text)  # :182
    timestep_proj = linear(F.silu(temb), sd, p + ".time_proj")  # :183
    txt = linear(F.gelu(linear(text, sd, p + ".text_embedder.linear_1"), approximate="tanh"), sd,
                 p + ".text_embedder.linear_2")  # :185, embeddings.py:2269-2273
    return temb, timestep_proj, txt


def wan_forward(sd: Dict[str, torch.Tensor], cfg: WanConfig, hidden_states: torch.Tensor, timestep: torch.Tensor,
                encoder_hidden_states: torch.Tensor, taps: Optional[dict] = None,
                num_layers: Optional[int] = None) -> torch.Tensor:
    """WanTransformer3DModel.forward (transformer_wan.py:454-552); returns ``sample`` [B, C_out, F, H, W]."""
    b, c, nf, hh, ww = hidden_states.shape
    p_t, p_h, p_w = cfg.patch_size
    ppf, pph, ppw = nf // p_t, hh // p_h, ww // p_w
    rotary = wan_rope(cfg, nf, hh, ww)  # :484
    x = F.conv3d(hidden_states, sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=cfg.patch_size)  # :486
    x = x.flatten(2).transpose(1, 2)  # :487
    if timestep.ndim == 2:  # :490-494
        ts_seq_len = timestep.shape[1]
        timestep = timestep.flatten()
    else:
        ts_seq_len = None
    temb, tproj, text = wan_condition_embedder(sd, cfg, timestep, encoder_hidden_states, ts_seq_len)  # :496
    tproj = tproj.unflatten(2, (6, -1)) if ts_seq_len is not None else tproj.unflatten(1, (6, -1))  # :499-504
    if taps is not None:
        taps["patch_embed"] = x
        taps["temb"] = temb
        taps["text"] = text
    n_layers = cfg.num_layers if num_layers is None else num_layers
    for i in range(n_layers):  # :516-517
        x = wan_block(sd, i, cfg, x, text, tproj, rotary, taps)
    sst = sd["scale_shift_table"]
    if temb.ndim == 3:  # :520-524
        shift, scale = (sst.unsqueeze(0) + temb.unsqueeze(2)).chunk(2, dim=2)
        shift, scale = shift.squeeze(2), scale.squeeze(2)
    else:  # :525-527
        shift, scale = (sst + temb.unsqueeze(1)).chunk(2, dim=1)
    x = (fp32_layer_norm(x.float(), None, None, cfg.eps) * (1 + scale) + shift).type_as(x)  # :536
    x = linear(x, sd, "proj_out")  # :537
    if taps is not None:
        taps["proj_out"] = x
    x = x.reshape(b, ppf, pph, ppw, p_t, p_h, p_w, -1)  # :539-541
    x = x.permute(0, 7, 1, 4, 2, 5, 3, 6)  # :542
    return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)  # :543
