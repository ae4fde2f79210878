"""CPU oracle (test infrastructure) — Wan VAE encode / decode, restated from
``/root/reference/architecture/autoencoder_kl_wan.py`` as pure functions over a diffusers-layout state dict.

The reference runs its encoder / decoder CHUNK BY CHUNK (one latent frame at a time in ``_decode`` :1198-1228, 1 + 4k
frames in ``_encode`` :1145-1170) and threads the last two input frames of every causal convolution through
``feat_cache``. This restatement is deliberately written the other way round — every layer over the WHOLE frame sequence
at once, causal convolutions as plain front-padded convolutions — so that agreeing with the reference's own chunked
execution (tests/golden/vae_golden.pt, produced by executing the reference file) also pins the cache bookkeeping:

  * a causal conv with the two cached frames prepended == a conv over the full sequence padded with two zero frames;
  * ``upsample3d`` (:265-295): the first latent frame bypasses ``time_conv`` (the ``"Rep"`` marker) and never enters its
    cache, so ``time_conv`` is a causal conv over frames 1.. with zero history; channel halves interleave as frames;
  * ``downsample3d`` (:301-311): frame 0 bypasses ``time_conv``; afterwards out_k = time_conv(g_{2k-2}, g_{2k-1}, g_{2k}),
    i.e. a valid stride-2 conv over the whole sequence, prefixed by g_0;
  * ``DupUp3D(first_chunk)`` (:109-131) duplicates every frame ``factor_t`` times and drops the first ``factor_t - 1``;
  * ``AvgDown3D`` (:55-87) front-pads the (odd-length) sequence with one zero frame, chunked or not.

Only the Wan2.2 form (``is_residual=True``: WanResidualDownBlock / WanResidualUpBlock) is restated — the FrameINO Wan2.2-5B
pipeline's VAE. Pinned: tests/test_oracle_golden.py compares with outputs of the reference classes (fp32, exact to 1e-5).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def causal_conv3d(x: torch.Tensor, sd, prefix: str, stride=(1, 1, 1), history: Optional[torch.Tensor] = None):
    """WanCausalConv3d (:134-176) over a whole sequence: time is padded at the FRONT with 2*pad_t zero frames (or the
    given history), space symmetrically."""
    w, b = sd[prefix + ".weight"], sd.get(prefix + ".bias")
    kt, kh, kw = w.shape[2:]
    pt, ph, pw = (kt - 1) // 2, (kh - 1) // 2, (kw - 1) // 2
    x = F.pad(x, (pw, pw, ph, ph, 2 * pt, 0))
    return F.conv3d(x, w, b, stride=stride)


def rms_norm(x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """WanRMS_norm (:179-202), channel_first: F.normalize over dim 1 * sqrt(C) * gamma."""
    g = gamma.reshape(1, -1, *([1] * (x.dim() - 2)))
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * g


def residual_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """WanResidualBlock.forward (:342-382)."""
    h = causal_conv3d(x, sd, p + ".conv_shortcut") if (p + ".conv_shortcut.weight") in sd else x
    x = F.silu(rms_norm(x, sd[p + ".norm1.gamma"]))
    x = causal_conv3d(x, sd, p + ".conv1")
    x = F.silu(rms_norm(x, sd[p + ".norm2.gamma"]))
    x = causal_conv3d(x, sd, p + ".conv2")
    return x + h


def attention_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """WanAttentionBlock.forward (:402-427): single-head attention over the pixels of each frame."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[p + ".norm.gamma"])
    qkv = F.conv2d(y, sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"])
    qkv = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous()
    q, k, v = qkv.chunk(3, dim=-1)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    o = F.conv2d(o, sd[p + ".proj.weight"], sd[p + ".proj.bias"])
    return o.view(b, t, c, h, w).permute(0, 2, 1, 3, 4) + x


def mid_block(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """WanMidBlock.forward (:455-466), num_layers = 1."""
    x = residual_block(sd, p + ".resnets.0", x)
    x = attention_block(sd, p + ".attentions.0", x)
    return residual_block(sd, p + ".resnets.1", x)


def _per_frame(x: torch.Tensor, fn):
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.view(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def resample(sd, p: str, mode: str, x: torch.Tensor) -> torch.Tensor:
    """WanResample.forward (:265-311) over the whole sequence (see the module docstring)."""
    if mode == "upsample3d" and x.shape[2] > 1:
        first, rest = x[:, :, :1], x[:, :, 1:]
        b, c, t, h, w = rest.shape
        y = causal_conv3d(rest, sd, p + ".time_conv")  # [b, 2c, t, h, w], zero history
        y = y.reshape(b, 2, c, t, h, w)
        y = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, t * 2, h, w)  # :291-293
        x = torch.cat([first, y], dim=2)
    if mode in ("upsample2d", "upsample3d"):
        def up(f):
            f = F.interpolate(f.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(f)  # WanUpsample
            return F.conv2d(f, sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], padding=1)
        return _per_frame(x, up)
    if mode in ("downsample2d", "downsample3d"):
        def down(f):
            return F.conv2d(F.pad(f, (0, 1, 0, 1)), sd[p + ".resample.1.weight"], sd[p + ".resample.1.bias"], stride=2)
        x = _per_frame(x, down)
        if mode == "downsample3d" and x.shape[2] > 1:
            w, bias = sd[p + ".time_conv.weight"], sd[p + ".time_conv.bias"]
            x = torch.cat([x[:, :, :1], F.conv3d(x, w, bias, stride=(2, 1, 1))], dim=2)  # :301-311
        return x
    return x


def dup_up3d(x: torch.Tensor, out_channels: int, ft: int, fs: int) -> torch.Tensor:
    """DupUp3D.forward (:109-131) on a whole sequence (first_chunk semantics: drop the first ft - 1 frames)."""
    repeats = out_channels * ft * fs * fs // x.shape[1]
    x = x.repeat_interleave(repeats, dim=1)
    x = x.view(x.size(0), out_channels, ft, fs, fs, x.size(2), x.size(3), x.size(4))
    x = x.permute(0, 1, 5, 2, 6, 3, 7, 4).contiguous()
    x = x.view(x.size(0), out_channels, x.size(2) * ft, x.size(4) * fs, x.size(6) * fs)
    return x[:, :, ft - 1:]


def avg_down3d(x: torch.Tensor, out_channels: int, ft: int, fs: int) -> torch.Tensor:
    """AvgDown3D.forward (:55-87)."""
    pad_t = (ft - x.shape[2] % ft) % ft
    x = F.pad(x, (0, 0, 0, 0, pad_t, 0))
    b, c, t, h, w = x.shape
    group = c * ft * fs * fs // out_channels
    x = x.view(b, c, t // ft, ft, h // fs, fs, w // fs, fs).permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    x = x.view(b, out_channels, group, t // ft, h // fs, w // fs)
    return x.mean(dim=2)


def _patchify(x: torch.Tensor, ps: Optional[int]) -> torch.Tensor:
    """:912-932"""
    if not ps or ps == 1:
        return x
    b, c, f, h, w = x.shape
    x = x.view(b, c, f, h // ps, ps, w // ps, ps).permute(0, 1, 6, 4, 2, 3, 5).contiguous()
    return x.view(b, c * ps * ps, f, h // ps, w // ps)


def _unpatchify(x: torch.Tensor, ps: Optional[int]) -> torch.Tensor:
    """:935-952"""
    if not ps or ps == 1:
        return x
    b, cp, f, h, w = x.shape
    c = cp // (ps * ps)
    x = x.view(b, c, ps, ps, f, h, w).permute(0, 1, 4, 5, 3, 6, 2).contiguous()
    return x.view(b, c, f, h * ps, w * ps)


def _dims(cfg: dict, decoder: bool) -> List[int]:
    mult = list(cfg["dim_mult"])
    if decoder:
        dim = cfg.get("decoder_base_dim") or cfg["base_dim"]
        return [dim * u for u in [mult[-1]] + mult[::-1]]  # :821
    return [cfg["base_dim"] * u for u in [1] + mult]  # :545


def decode(sd: Dict[str, torch.Tensor], cfg: dict, z: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """AutoencoderKLWan._decode (:1198-1228) + WanDecoder3d.forward (:874-909): z [B, z_dim, T, h, w] -> video
    [B, 3, 1 + 4 (T - 1), 16 h, 16 w] clamped to [-1, 1]."""
    assert cfg.get("is_residual", False), "only the Wan2.2 (is_residual) VAE is restated"
    x = causal_conv3d(z, sd, "post_quant_conv")  # :1207
    x = causal_conv3d(x, sd, "decoder.conv_in")
    x = mid_block(sd, "decoder.mid_block", x)
    if taps is not None:
        taps["mid"] = x
    dims = _dims(cfg, True)
    t_up = list(cfg["temperal_downsample"])[::-1]  # :1036
    n = len(cfg["dim_mult"])
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        p = f"decoder.up_blocks.{i}"
        up_flag = i != n - 1
        x_copy = x
        for j in range(cfg["num_res_blocks"] + 1):
            x = residual_block(sd, f"{p}.resnets.{j}", x)
        if up_flag:
            x = resample(sd, p + ".upsampler", "upsample3d" if t_up[i] else "upsample2d", x)
            x = x + dup_up3d(x_copy, out_dim, 2 if t_up[i] else 1, 2)  # :709-710
        if taps is not None:
            taps[f"up{i}"] = x
    x = F.silu(rms_norm(x, sd["decoder.norm_out.gamma"]))
    x = causal_conv3d(x, sd, "decoder.conv_out")
    if taps is not None:
        taps["head"] = x  # before un-patchify and clamp
    x = _unpatchify(x, cfg.get("patch_size"))  # :1221-1222
    return torch.clamp(x, min=-1.0, max=1.0)  # :1224


def encode(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """AutoencoderKLWan._encode (:1145-1170) + WanEncoder3d.forward (:586-623): video [B, 3, 1 + 4k, H, W] -> the
    posterior parameters [B, 2 z_dim, 1 + k, H/16, W/16] (mean = first z_dim channels: DiagonalGaussianDistribution.mode)."""
    assert cfg.get("is_residual", False), "only the Wan2.2 (is_residual) VAE is restated"
    x = _patchify(x, cfg.get("patch_size"))  # :1152-1153
    x = causal_conv3d(x, sd, "encoder.conv_in")
    dims = _dims(cfg, False)
    t_down = list(cfg["temperal_downsample"])
    n = len(cfg["dim_mult"])
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        p = f"encoder.down_blocks.{i}"
        down_flag = i != n - 1
        t_flag = t_down[i] if down_flag else False
        x_copy = x
        for j in range(cfg["num_res_blocks"]):
            x = residual_block(sd, f"{p}.resnets.{j}", x)
        if down_flag:
            x = resample(sd, p + ".downsampler", "downsample3d" if t_flag else "downsample2d", x)
        x = x + avg_down3d(x_copy, out_dim, 2 if t_flag else 1, 2 if down_flag else 1)  # :502
        if taps is not None:
            taps[f"down{i}"] = x
    x = mid_block(sd, "encoder.mid_block", x)
    x = F.silu(rms_norm(x, sd["encoder.norm_out.gamma"]))
    x = causal_conv3d(x, sd, "encoder.conv_out")
    return causal_conv3d(x, sd, "quant_conv")  # :1168
