"""Tensor-level wrappers over the C ABI (``include/frameino_b200.h``).

PyTorch is used here only for device memory and the current CUDA stream; all arithmetic happens in the hand-written
sm_100a kernels of ``frameino_b200/csrc``. Every function raises if its inputs are not CUDA tensors — there is no CPU
path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

EPI_NONE, EPI_GELU_TANH, EPI_SILU, EPI_GATE_RESIDUAL = 0, 1, 2, 3
GEMM_FLAG_ROUND_PRODUCT = 1
LN_FLAG_BF16_STEPS = 1
QK_RMS_ACROSS_HEADS, QK_LAYERNORM_PER_HEAD = 0, 1
ROPE_NONE, ROPE_WAN, ROPE_COGVIDEOX = 0, 1, 2

_bound_device = None


def _prep(*tensors: Optional[torch.Tensor]):
    """Validates devices, binds the library to the device once, returns (lib, stream pointer)."""
    global _bound_device
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("frameino_b200 ops need CUDA tensors (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
    lib = _lib.load()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if _bound_device != idx:
        _lib.check(lib.fino_set_device(idx), "fino_set_device")
        _bound_device = idx
    return lib, torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows2d(t: torch.Tensor) -> Tuple[int, int, int]:
    """(rows, cols, row_stride) of a tensor viewed as a 2-D row matrix with contiguous columns (no copy)."""
    if t.dim() < 2:
        raise ValueError("need at least 2 dims")
    if t.stride(-1) != 1:
        raise ValueError("last dimension must be contiguous")
    t2 = t if t.dim() == 2 else t.view(-1, t.shape[-1])  # raises if the leading dims do not collapse
    return t2.shape[0], t2.shape[1], (t2.stride(0) if t2.shape[0] > 1 else t2.shape[1])


def linear(
    x: torch.Tensor,
    weight: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    *,
    epilogue: int = EPI_NONE,
    out: Optional[torch.Tensor] = None,
    out_dtype: torch.dtype = torch.bfloat16,
    residual: Optional[torch.Tensor] = None,
    gate: Optional[torch.Tensor] = None,
    row_index: Optional[torch.Tensor] = None,
    rows_per_group: int = 0,
    round_product: bool = False,
) -> torch.Tensor:
    """``epilogue(x @ weight.T + bias)`` on the tcgen05 GEMM. x: [..., K] bf16, weight: [N, K] bf16.

    With ``EPI_GATE_RESIDUAL``: ``residual + bf16(x @ W.T + b) * gate[row_index[row]]`` (gate fp32 [R, >=N]).
    """
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
    lib, stream = _prep(x, weight, bias, out, residual, gate, row_index)
    m, k, lda = _rows2d(x)
    n, k2 = weight.shape
    assert k2 == k, f"K mismatch {k2} vs {k}"
    assert weight.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.is_contiguous() and bias.numel() == n
    if out is None:
        out = torch.empty(*x.shape[:-1], n, dtype=out_dtype, device=x.device)
    om, on, ldc = _rows2d(out)
    assert om == m and on == n
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.bfloat16
        rm, rn, ldr = _rows2d(residual)
        assert rm == m and rn == n
    gstride = 0
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.stride(-1) == 1
        gstride = gate.stride(0) if gate.dim() == 2 else 0
    if row_index is not None:
        assert row_index.dtype == torch.int32 and row_index.is_contiguous() and row_index.numel() == m
    status = lib.fino_gemm_bf16(
        x.data_ptr(), lda, weight.data_ptr(), weight.stride(0), _ptr(bias), out.data_ptr(), ldc, m, n, k, epilogue,
        1 if out.dtype == torch.float32 else 0, GEMM_FLAG_ROUND_PRODUCT if round_product else 0, _ptr(residual), ldr,
        _ptr(gate), gstride, _ptr(row_index), rows_per_group, stream,
    )
    _lib.check(status, "fino_gemm_bf16")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Non-causal attention over token-major tensors. q: [B, Nq, H*d], k/v: [B, Nk, H*d] (any row/batch strides)."""
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.dim() == 3 and k.dim() == 3 and v.dim() == 3
    lib, stream = _prep(q, k, v, out)
    b, nq, inner = q.shape
    nk = k.shape[1]
    d = inner // heads
    assert d * heads == inner and k.shape == (b, nk, inner) and v.shape == (b, nk, inner)
    for t in (q, k, v):
        assert t.stride(2) == 1
    if scale is None:
        scale = d ** -0.5
    if d not in (64, 128):
        # The tcgen05 kernel is built for the production head dims (128 Wan, 64 CogVideoX). Other head dims (the tiny
        # test config uses 32) are zero-padded per head to the next supported width: zero columns add nothing to
        # Q K^T and the padded output columns are dropped, so the result is unchanged.
        if d > 128 or d % 8 != 0:
            raise _lib.FinoError(f"attention: head_dim {d} unsupported (multiples of 8 up to 128)")
        dp = 64 if d < 64 else 128

        def pad(t):
            n = t.shape[1]
            tp = torch.zeros(b, n, heads, dp, dtype=t.dtype, device=t.device)
            tp[..., :d] = t.reshape(b, n, heads, d) if t.is_contiguous() else t.unflatten(2, (heads, d))
            return tp.view(b, n, heads * dp)

        op = attention(pad(q), pad(k), pad(v), heads, scale=scale)
        res = op.view(b, nq, heads, dp)[..., :d].reshape(b, nq, inner)
        if out is not None:
            out.copy_(res)
            return out
        return res
    if out is None:
        out = torch.empty(b, nq, inner, dtype=torch.bfloat16, device=q.device)
    assert out.shape == (b, nq, inner) and out.stride(2) == 1
    status = lib.fino_attention_fwd(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, heads, nq, nk, d,
        q.stride(1), k.stride(1), v.stride(1), out.stride(1), q.stride(0), k.stride(0), v.stride(0), out.stride(0),
        float(scale), stream,
    )
    _lib.check(status, "fino_attention_fwd")
    return out


def ln_modulate(
    x: torch.Tensor,
    eps: float,
    *,
    gamma: Optional[torch.Tensor] = None,
    beta: Optional[torch.Tensor] = None,
    shift: Optional[torch.Tensor] = None,
    scale: Optional[torch.Tensor] = None,
    row_index: Optional[torch.Tensor] = None,
    rows_per_group: int = 0,
    bf16_steps: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """LayerNorm(x) [*gamma+beta] [*(1+scale)+shift]. shift/scale: fp32 [R, dim] views sharing one row stride."""
    assert x.dtype == torch.bfloat16
    lib, stream = _prep(x, gamma, beta, shift, scale, row_index, out)
    rows, dim, xs = _rows2d(x)
    if out is None:
        out = torch.empty_like(x)
    _, _, os_ = _rows2d(out)
    mstride = 0
    if shift is not None:
        assert scale is not None and shift.dtype == scale.dtype == torch.float32
        assert shift.dim() == 2 and scale.dim() == 2 and shift.stride(1) == 1 and scale.stride(1) == 1
        assert shift.stride(0) == scale.stride(0) and shift.shape[1] == dim
        mstride = shift.stride(0)
    for t in (gamma, beta):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == dim
    if row_index is not None:
        assert row_index.dtype == torch.int32 and row_index.numel() == rows
    status = lib.fino_ln_modulate(x.data_ptr(), out.data_ptr(), rows, dim, xs, os_, float(eps), _ptr(gamma),
                                  _ptr(beta), _ptr(shift), _ptr(scale), mstride, _ptr(row_index), rows_per_group,
                                  LN_FLAG_BF16_STEPS if bf16_steps else 0, stream)
    _lib.check(status, "fino_ln_modulate")
    return out


def gate_residual(x: torch.Tensor, y: torch.Tensor, gate: Optional[torch.Tensor] = None,
                  row_index: Optional[torch.Tensor] = None, rows_per_group: int = 0, round_product: bool = False,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert x.dtype == y.dtype == torch.bfloat16
    lib, stream = _prep(x, y, gate, row_index, out)
    rows, dim, xs = _rows2d(x)
    r2, d2, ys = _rows2d(y)
    assert (rows, dim) == (r2, d2)
    if out is None:
        out = torch.empty_like(x)
    _, _, os_ = _rows2d(out)
    gs = 0
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.dim() == 2 and gate.stride(1) == 1
        gs = gate.stride(0)
    status = lib.fino_gate_residual(x.data_ptr(), y.data_ptr(), out.data_ptr(), rows, dim, xs, ys, os_, _ptr(gate), gs,
                                    _ptr(row_index), rows_per_group, 1 if round_product else 0, stream)
    _lib.check(status, "fino_gate_residual")
    return out


def qk_norm_rope(
    x0: torch.Tensor,
    w0: Optional[torch.Tensor],
    x1: Optional[torch.Tensor],
    w1: Optional[torch.Tensor],
    heads: int,
    *,
    b0: Optional[torch.Tensor] = None,
    b1: Optional[torch.Tensor] = None,
    rope0: bool = True,
    rope1: bool = True,
    norm_mode: int = QK_RMS_ACROSS_HEADS,
    eps: float = 1e-6,
    rope_mode: int = ROPE_NONE,
    cos: Optional[torch.Tensor] = None,
    sin: Optional[torch.Tensor] = None,
    seq_len: int = 0,
    rope_skip: int = 0,
) -> None:
    """In-place q/k norm + RoPE. x0/x1: [..., H*d] row views (e.g. column slices of a fused QKV buffer)."""
    lib, stream = _prep(x0, w0, x1, w1, b0, b1, cos, sin)
    rows0, dim, s0 = _rows2d(x0)
    head_dim = dim // heads
    rows1, s1 = 0, 0
    if x1 is not None:
        rows1, dim1, s1 = _rows2d(x1)
        assert dim1 == dim
    for t in (w0, w1, b0, b1):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.is_contiguous()
    if cos is not None:
        assert cos.dtype == sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape[-1] == head_dim
    status = lib.fino_qk_norm_rope(x0.data_ptr(), rows0, s0, _ptr(w0), _ptr(b0), 1 if rope0 else 0, _ptr(x1), rows1,
                                   s1, _ptr(w1), _ptr(b1), 1 if rope1 else 0, heads, head_dim, norm_mode, float(eps),
                                   rope_mode, _ptr(cos), _ptr(sin), seq_len, rope_skip, stream)
    _lib.check(status, "fino_qk_norm_rope")


def patchify(x: torch.Tensor, dims: Tuple[int, int, int, int, int], strides: Tuple[int, int, int, int, int],
             patch: Tuple[int, int, int]) -> torch.Tensor:
    """x viewed as [B,C,F,H,W] through (dims, element strides) -> rows [B*F/pt*H/ph*W/pw, C*pt*ph*pw]."""
    assert x.dtype == torch.bfloat16
    lib, stream = _prep(x)
    b, c, f, h, w = dims
    pt, ph, pw = patch
    rows = b * (f // pt) * (h // ph) * (w // pw)
    kdim = c * pt * ph * pw
    out = torch.empty(rows, kdim, dtype=torch.bfloat16, device=x.device)
    status = lib.fino_patchify(x.data_ptr(), out.data_ptr(), b, c, f, h, w, pt, ph, pw, *strides, kdim, stream)
    _lib.check(status, "fino_patchify")
    return out


def unpatchify(rows: torch.Tensor, out: torch.Tensor, dims: Tuple[int, int, int, int, int],
               strides: Tuple[int, int, int, int, int], patch: Tuple[int, int, int], channel_last: bool) -> torch.Tensor:
    assert rows.dtype == out.dtype == torch.bfloat16 and rows.dim() == 2 and rows.stride(1) == 1
    lib, stream = _prep(rows, out)
    b, c, f, h, w = dims
    pt, ph, pw = patch
    status = lib.fino_unpatchify(rows.data_ptr(), out.data_ptr(), b, c, f, h, w, pt, ph, pw, *strides, rows.stride(0),
                                 1 if channel_last else 0, stream)
    _lib.check(status, "fino_unpatchify")
    return out


def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0.0,
                       scale: float = 1.0, max_period: float = 10000.0) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.dim() == 1 and t.is_contiguous()
    lib, stream = _prep(t)
    out = torch.empty(t.numel(), dim, dtype=torch.float32, device=t.device)
    status = lib.fino_timestep_embedding(t.data_ptr(), out.data_ptr(), t.numel(), dim, 1 if flip_sin_to_cos else 0,
                                         float(downscale_freq_shift), float(scale), float(max_period), stream)
    _lib.check(status, "fino_timestep_embedding")
    return out


def linear_small_m(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], *, act_in: int = 0,
                   act_out: int = 0, round_in: bool = False, round_out: bool = False) -> torch.Tensor:
    """fp32 [m, K] x (fp32|bf16) [N, K] -> fp32 [m, N] with fp32 accumulation (built for m <= 8; larger m runs as
    independent 8-row chunks inside one launch)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    assert weight.dtype in (torch.float32, torch.bfloat16) and weight.is_contiguous()
    if bias is not None:
        assert bias.dtype == weight.dtype and bias.is_contiguous()
    lib, stream = _prep(x, weight, bias)
    m, k = x.shape
    n = weight.shape[0]
    assert weight.shape[1] == k
    out = torch.empty(m, n, dtype=torch.float32, device=x.device)
    status = lib.fino_linear_small_m(x.data_ptr(), weight.data_ptr(), _ptr(bias), out.data_ptr(), m, n, k,
                                     1 if weight.dtype == torch.bfloat16 else 0, act_in, act_out,
                                     1 if round_in else 0, 1 if round_out else 0, stream)
    _lib.check(status, "fino_linear_small_m")
    return out


def timestep_dedup(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Device-side de-duplication of per-token timesteps (no host sync): fp32 [n] -> (uniq fp32 [8] ascending with the
    unused slots repeating the largest value, row_index int32 [n], count int32 [1] = distinct values, 9 = more than 8)."""
    assert t.dtype == torch.float32 and t.dim() == 1 and t.is_contiguous() and t.numel() > 0
    lib, stream = _prep(t)
    uniq = torch.empty(8, dtype=torch.float32, device=t.device)
    row_index = torch.empty(t.numel(), dtype=torch.int32, device=t.device)
    count = torch.empty(1, dtype=torch.int32, device=t.device)
    _lib.check(lib.fino_timestep_dedup(t.data_ptr(), t.numel(), uniq.data_ptr(), row_index.data_ptr(), count.data_ptr(),
                                       stream), "fino_timestep_dedup")
    return uniq, row_index, count


def build_mod_table(table: torch.Tensor, proj: torch.Tensor, layers: int, table_layer_stride: int) -> torch.Tensor:
    """out[l, r, :] = table[l*stride : l*stride+cols] + proj[r, :] (fp32)."""
    assert table.dtype == proj.dtype == torch.float32 and proj.dim() == 2 and proj.is_contiguous()
    lib, stream = _prep(table, proj)
    r, cols = proj.shape
    out = torch.empty(layers, r, cols, dtype=torch.float32, device=proj.device)
    status = lib.fino_build_mod_table(table.data_ptr(), proj.data_ptr(), out.data_ptr(), layers, r, cols,
                                      table_layer_stride, stream)
    _lib.check(status, "fino_build_mod_table")
    return out


def _f32c(t: Optional[torch.Tensor], shape, what: str) -> None:
    if t is None:
        return
    if t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
        raise ValueError(f"{what}: want contiguous float32 {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")


def wan_pack_model_input(latents: torch.Tensor, condition: torch.Tensor, mask: torch.Tensor,
                         id_latents: Optional[torch.Tensor], traj_latents: torch.Tensor,
                         patch: Tuple[int, int, int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Blend + ID concat + trajectory concat + bf16 cast + patchify of one sampler step (``fino_wan_pack_model_input``).
    latents/condition: fp32 [B, C, F, H, W]; mask: fp32 [F, H, W]; id_latents: fp32 [B, C, n_id, H, W] or None;
    traj_latents: fp32 [B, C, F + n_id, H, W]. Returns bf16 rows [B * tokens, 2*C*pt*ph*pw]."""
    b, c, f, h, w = latents.shape
    n_id = 0 if id_latents is None else id_latents.shape[2]
    _f32c(latents, (b, c, f, h, w), "latents")
    _f32c(condition, (b, c, f, h, w), "condition")
    _f32c(mask, (f, h, w), "mask")
    _f32c(id_latents, (b, c, n_id, h, w), "id_latents")
    _f32c(traj_latents, (b, c, f + n_id, h, w), "traj_latents")
    lib, stream = _prep(latents, condition, mask, id_latents, traj_latents, out)
    pt, ph, pw = patch
    rows = b * ((f + n_id) // pt) * (h // ph) * (w // pw)
    kdim = 2 * c * pt * ph * pw
    if out is None:
        out = torch.empty(rows, kdim, dtype=torch.bfloat16, device=latents.device)
    assert out.dtype == torch.bfloat16 and out.shape == (rows, kdim) and out.stride(1) == 1
    status = lib.fino_wan_pack_model_input(latents.data_ptr(), condition.data_ptr(), mask.data_ptr(), _ptr(id_latents),
                                           traj_latents.data_ptr(), out.data_ptr(), b, c, f, n_id, h, w, pt, ph, pw,
                                           out.stride(0), stream)
    _lib.check(status, "fino_wan_pack_model_input")
    return out


def wan_cfg_euler_step(latents: torch.Tensor, y_cond: torch.Tensor, y_uncond: Optional[torch.Tensor], n_id: int,
                       patch: Tuple[int, int, int], guidance: float, dsigma: float) -> torch.Tensor:
    """In place: latents += dsigma * (y_uncond + guidance * (y_cond - y_uncond)) from the forwards' proj_out rows
    (``fino_wan_cfg_euler_step``; CFG + ID-frame drop + Euler step + un-patchify). latents: fp32 [B, C, F, H, W];
    y_*: bf16 [B * tokens(F + n_id), pt*ph*pw*C]."""
    b, c, f, h, w = latents.shape
    _f32c(latents, (b, c, f, h, w), "latents")
    lib, stream = _prep(latents, y_cond, y_uncond)
    pt, ph, pw = patch
    rows = b * ((f + n_id) // pt) * (h // ph) * (w // pw)
    for y in (y_cond, y_uncond):
        if y is not None:
            assert y.dtype == torch.bfloat16 and y.dim() == 2 and y.stride(1) == 1
            assert y.shape == (rows, c * pt * ph * pw), f"rows {tuple(y.shape)} != {(rows, c * pt * ph * pw)}"
    if y_uncond is not None:
        assert y_uncond.stride(0) == y_cond.stride(0)
    status = lib.fino_wan_cfg_euler_step(y_cond.data_ptr(), _ptr(y_uncond), y_cond.stride(0), latents.data_ptr(), b, c,
                                         f, n_id, h, w, pt, ph, pw, float(guidance), float(dsigma), stream)
    _lib.check(status, "fino_wan_cfg_euler_step")
    return latents


# ---------------------------------------------------------------------------------------------------------------
# Wan VAE (include/frameino_b200.h "Wan VAE" section): channels-last bf16 activations [T, H, W, C]
# ---------------------------------------------------------------------------------------------------------------
def conv3d_cl(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], kernel: Tuple[int, int, int],
              out: Optional[torch.Tensor] = None, pad_hw: Tuple[int, int] = (0, 0), stride_hw: int = 1, stride_t: int = 1,
              residual: Optional[torch.Tensor] = None, out_hw: Optional[Tuple[int, int]] = None,
              c_out: Optional[int] = None) -> torch.Tensor:
    """Implicit-GEMM convolution, valid along time (prepend the causal history). x: bf16 [T_in, H, W, C_in] (pixel rows
    contiguous, any frame / row / pixel strides); weight: packed bf16 [C_out, kt*kh*kw*ceil(C_in/64)*64]; out: bf16
    [T_out, H_out, W_out, C_out] view (frame / row strides multiples of the pixel stride); residual: same geometry."""
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() == 4 and x.stride(3) == 1
    lib, stream = _prep(x, weight, bias, out, residual)
    kt, kh, kw = kernel
    t_in, h_in, w_in, c_in = x.shape
    t_out = (t_in - kt) // stride_t + 1
    if out_hw is None:
        h_out = (h_in + 2 * pad_hw[0] - kh) // stride_hw + 1
        w_out = (w_in + 2 * pad_hw[1] - kw) // stride_hw + 1
    else:
        h_out, w_out = out_hw
    n = weight.shape[0] if c_out is None else c_out
    if out is None:
        out = torch.empty(t_out, h_out, w_out, n, dtype=torch.bfloat16, device=x.device)
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == (t_out, h_out, w_out, n) and out.stride(3) == 1
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.shape == out.shape and residual.stride() == out.stride()
    if bias is not None:
        assert bias.dtype == torch.bfloat16 and bias.is_contiguous() and bias.numel() >= n
    status = lib.fino_conv3d_cl_bf16(
        x.data_ptr(), t_in, h_in, w_in, c_in, x.stride(0), x.stride(1), x.stride(2), weight.data_ptr(), weight.stride(0),
        _ptr(bias), out.data_ptr(), t_out, h_out, w_out, n, out.stride(0), out.stride(1), out.stride(2), kt, kh, kw,
        pad_hw[0], pad_hw[1], stride_hw, stride_t, _ptr(residual), EPI_GATE_RESIDUAL if residual is not None else EPI_NONE,
        stream)
    _lib.check(status, "fino_conv3d_cl_bf16")
    return out


def rms_act_cl(x: torch.Tensor, gamma: torch.Tensor, bias: Optional[torch.Tensor] = None, silu: bool = True,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """WanRMS_norm over the channels of every pixel (+ SiLU). x/out: bf16 [..., C] with contiguous pixels."""
    assert x.dtype == torch.bfloat16 and gamma.dtype == torch.float32 and gamma.is_contiguous()
    lib, stream = _prep(x, gamma, bias, out)
    rows, c, xs = _rows2d(x)
    if out is None:
        out = torch.empty_like(x)
    orows, oc, os_ = _rows2d(out)
    assert (orows, oc) == (rows, c) and gamma.numel() == c
    _lib.check(lib.fino_rms_act_cl(x.data_ptr(), out.data_ptr(), rows, c, xs, os_, gamma.data_ptr(), _ptr(bias),
                                   1 if silu else 0, stream), "fino_rms_act_cl")
    return out


def upsample2x_cl(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
    lib, stream = _prep(x, out)
    t, h, w, c = x.shape
    if out is None:
        out = torch.empty(t, 2 * h, 2 * w, c, dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype and tuple(out.shape) == (t, 2 * h, 2 * w, c) and (t == 1 or out.is_contiguous())
    assert out[0].is_contiguous()
    _lib.check(lib.fino_upsample2x_cl(x.data_ptr(), out.data_ptr(), t, h, w, c, stream), "fino_upsample2x_cl")
    return out


def dupup_add_cl(y: torch.Tensor, src: torch.Tensor, ft: int, fs: int, first_chunk: bool) -> torch.Tensor:
    assert y.dtype == src.dtype == torch.bfloat16 and y.is_contiguous() and src.is_contiguous()
    lib, stream = _prep(y, src)
    _lib.check(lib.fino_dupup_add_cl(y.data_ptr(), src.data_ptr(), *y.shape, *src.shape, ft, fs,
                                     ft - 1 if first_chunk else 0, stream), "fino_dupup_add_cl")
    return y


def avgdown_add_cl(y: torch.Tensor, src: torch.Tensor, ft: int, fs: int) -> torch.Tensor:
    assert y.dtype == src.dtype == torch.bfloat16 and y.is_contiguous() and src.is_contiguous()
    lib, stream = _prep(y, src)
    _lib.check(lib.fino_avgdown_add_cl(y.data_ptr(), src.data_ptr(), *y.shape, *src.shape, ft, fs, stream),
               "fino_avgdown_add_cl")
    return y


def softmax_rows(s: torch.Tensor, scale: float, out: torch.Tensor) -> torch.Tensor:
    """out[r, :cols] = softmax(s[r] * scale) as bf16; out's extra columns (row stride padding) are zeroed."""
    assert s.dtype == torch.float32 and s.dim() == 2 and s.stride(1) == 1 and out.dtype == torch.bfloat16
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[0] == s.shape[0] and out.shape[1] >= s.shape[1]
    lib, stream = _prep(s, out)
    _lib.check(lib.fino_softmax_rows(s.data_ptr(), out.data_ptr(), s.shape[0], s.shape[1], s.stride(0), out.stride(0),
                                     float(scale), stream), "fino_softmax_rows")
    return out


def vae_to_cl(x: torch.Tensor, patch: int, cpad: int) -> torch.Tensor:
    """[C, T, H, W] fp32 | bf16 (any strides) -> channels-last bf16 [T, H/patch, W/patch, cpad] (patchified, zero padded)."""
    assert x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16)
    lib, stream = _prep(x)
    c, t, h, w = x.shape
    out = torch.empty(t, h // patch, w // patch, cpad, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.fino_vae_to_cl(x.data_ptr(), 1 if x.dtype == torch.float32 else 0, out.data_ptr(), c, t, h, w,
                                  *x.stride(), patch, cpad, stream), "fino_vae_to_cl")
    return out


def vae_from_cl(x: torch.Tensor, out: torch.Tensor, channels: int, patch: int, clamp: bool) -> torch.Tensor:
    """channels-last bf16 [T, Hi, Wi, cstride] -> out[channels, T, Hi*patch, Wi*patch] (a [C, T, H, W] view whose frames,
    rows and pixels are contiguous; fp32 | bf16), un-patchified and optionally clamped to [-1, 1]."""
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous() and out.dtype in (torch.float32, torch.bfloat16)
    t, hi, wi, cs = x.shape
    assert tuple(out.shape) == (channels, t, hi * patch, wi * patch)
    assert out.stride(3) == 1 and out.stride(2) == wi * patch and out.stride(1) == hi * patch * wi * patch
    lib, stream = _prep(x, out)
    _lib.check(lib.fino_vae_from_cl(x.data_ptr(), out.data_ptr(), 1 if out.dtype == torch.float32 else 0, channels, t, hi,
                                    wi, cs, patch, 1 if clamp else 0, out.stride(0), stream), "fino_vae_from_cl")
    return out


def swap01(x: torch.Tensor, a: int, b: int, inner: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Contiguous [a, b, inner] -> [b, a, inner] (bf16, inner % 8 == 0)."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() == a * b * inner
    lib, stream = _prep(x, out)
    if out is None:
        out = torch.empty(b, a, inner, dtype=x.dtype, device=x.device)
    assert out.is_contiguous() and out.numel() == x.numel()
    _lib.check(lib.fino_swap01(x.data_ptr(), out.data_ptr(), a, b, inner, stream), "fino_swap01")
    return out


def gemm_set_mode(mode: int) -> None:
    """0 auto, 1 single-CTA tcgen05 GEMM, 2 CTA-pair (cta_group::2) GEMM."""
    _lib.check(_lib.load().fino_gemm_set_mode(mode), "fino_gemm_set_mode")


def gemm_set_split(mode: int) -> None:
    """Split-K of the pair GEMM's partly filled last round: -1 automatic (default), 0 never, 2..16 forced (test hook)."""
    _lib.check(_lib.load().fino_gemm_set_split(int(mode)), "fino_gemm_set_split")


def gemm_plan(m: int, n: int, k: int, sms: int = 148, mode: int = -1):
    """(num_full, splits) of the pair GEMM on a device with ``sms`` SMs (host arithmetic only)."""
    import ctypes

    num_full, splits = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().fino_gemm_plan(m, n, k, sms, mode, ctypes.byref(num_full), ctypes.byref(splits)),
               "fino_gemm_plan")
    return num_full.value, splits.value


def attention_set_variant(variant: int) -> None:
    _lib.check(_lib.load().fino_attention_set_variant(variant), "fino_attention_set_variant")


def attention_set_split(mode: int) -> None:
    """KV split of the attention kernel's partial last wave: -1 automatic (default), 0 never, 2..64 split every tile
    that many ways (test hook). See include/frameino_b200.h."""
    _lib.check(_lib.load().fino_attention_set_split(int(mode)), "fino_attention_set_split")


def attention_plan(nq: int, nk: int, heads: int, batch: int = 1, sms: int = 148, mode: int = -1):
    """(n_full, splits) the attention launcher would use on a device with ``sms`` SMs (host arithmetic only)."""
    import ctypes

    n_full, splits = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().fino_attention_plan(nq, nk, heads, batch, sms, mode, ctypes.byref(n_full),
                                               ctypes.byref(splits)), "fino_attention_plan")
    return n_full.value, splits.value


def attention_plan_hd(nq: int, nk: int, heads: int, head_dim: int, batch: int = 1, sms: int = 148, mode: int = -1):
    """(n_full, splits, tile_rows) for a head_dim: 128 -> 256-row tiles, 64 -> the four-tile kernel's 512-row tiles."""
    import ctypes

    n_full, splits, rows = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().fino_attention_plan_hd(nq, nk, heads, batch, head_dim, sms, mode, ctypes.byref(n_full),
                                                  ctypes.byref(splits), ctypes.byref(rows)), "fino_attention_plan_hd")
    return n_full.value, splits.value, rows.value


def rows_set_variant(ln_block: int, qk_block: int) -> None:
    """LayerNorm kernel: 0 warp per row, 1 block per row (one row at a time), 2 batched block per row (default for
    1024 <= dim <= 3072). q/k kernel: 0 warp per row, 1 block per token (default)."""
    _lib.check(_lib.load().fino_rows_set_variant(int(ln_block), int(qk_block)), "fino_rows_set_variant")


def rows_set_tma(on: bool) -> None:
    """Experimental TMA-staged persistent row kernels for wide rows (default off = register-resident row kernels)."""
    _lib.check(_lib.load().fino_rows_set_tma(int(on)), "fino_rows_set_tma")


def launch_count() -> int:
    return int(_lib.load().fino_launch_count())


# ---------------------------------------------------------------------------------------------------------------
# peer memory (Ulysses exchange over NVLink / NVSwitch, include/frameino_b200.h "peer" section)
# ---------------------------------------------------------------------------------------------------------------
PEER_MAX_RANKS = 8


def _bind_current_device():
    global _bound_device
    lib = _lib.load()
    idx = torch.cuda.current_device()
    if _bound_device != idx:
        _lib.check(lib.fino_set_device(idx), "fino_set_device")
        _bound_device = idx
    return lib


def peer_alloc(nbytes: int) -> int:
    """cudaMalloc'd, zero-filled, IPC-exportable buffer on the current device; returns the device pointer."""
    import ctypes

    lib = _bind_current_device()
    p = ctypes.c_void_p()
    _lib.check(lib.fino_peer_alloc(int(nbytes), ctypes.byref(p)), "fino_peer_alloc")
    return int(p.value)


def peer_free(ptr: int) -> None:
    _lib.check(_bind_current_device().fino_peer_free(ptr), "fino_peer_free")


def peer_export(ptr: int) -> bytes:
    import ctypes

    buf = ctypes.create_string_buffer(64)
    _lib.check(_bind_current_device().fino_peer_export(ptr, buf), "fino_peer_export")
    return bytes(buf.raw)


def peer_import(handle: bytes) -> int:
    import ctypes

    assert len(handle) == 64
    p = ctypes.c_void_p()
    _lib.check(_bind_current_device().fino_peer_import(handle, ctypes.byref(p)), "fino_peer_import")
    return int(p.value)


def peer_release(ptr: int) -> None:
    _lib.check(_bind_current_device().fino_peer_release(ptr), "fino_peer_release")


def pointer_table(ptrs) -> "ctypes.Array":
    """HOST array of up to 8 device pointers, the form every ``*_ptrs`` argument of the peer entry points takes."""
    import ctypes

    assert 1 <= len(ptrs) <= PEER_MAX_RANKS
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


class _RawCudaBytes:
    """__cuda_array_interface__ shim: lets torch view memory this library allocated (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def tensor_from_ptr(ptr: int, shape, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """Tensor view (current device) of raw device memory, e.g. a slice of a ``peer_alloc`` buffer."""
    numel = 1
    for s in shape:
        numel *= int(s)
    nbytes = numel * torch.empty(0, dtype=dtype).element_size()
    raw = torch.as_tensor(_RawCudaBytes(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
    return raw.view(dtype).view(*shape)


def peer_barrier(flag_ptrs, rank: int, world: int, epoch: int) -> None:
    """Stream-ordered barrier between the ranks through their peer-mapped flag arrays (``pointer_table``)."""
    lib = _bind_current_device()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.fino_peer_barrier(flag_ptrs, rank, world, epoch & 0xFFFFFFFF, stream), "fino_peer_barrier")


def peer_status(own_flags_ptr: int):
    """[8] ints: entry t != 0 means a peer barrier on this rank gave up waiting for rank t (see fino_peer_barrier).
    Synchronises the current stream."""
    import ctypes

    lib = _bind_current_device()
    out = (ctypes.c_uint32 * PEER_MAX_RANKS)()
    _lib.check(lib.fino_peer_status(own_flags_ptr, out, torch.cuda.current_stream().cuda_stream), "fino_peer_status")
    return list(out)


HALO_DATA_OFF = 256  # mailbox header bytes (fino_halo_exchange)


def halo_exchange(frames: torch.Tensor, own: int, up: Optional[int], down: Optional[int], seq: int, slot_bytes: int,
                  rank: int) -> None:
    """Row-parallel convolution halo rows through peer memory. frames: bf16 [t, hl + 2, W, C] with contiguous frames
    interior (rows 1..hl written); own / up / down: mailbox device pointers (up / down None at the image border)."""
    assert frames.dtype == torch.bfloat16 and frames.dim() == 4 and frames[0].is_contiguous()
    lib, stream = _prep(frames)
    t, hp, w, c = frames.shape
    fs = frames.stride(0) if t > 1 else hp * w * c
    _lib.check(lib.fino_halo_exchange(frames.data_ptr(), t, hp - 2, w * c * 2, fs * 2, up, down, own,
                                      seq & 0xFFFFFFFF, slot_bytes, rank, stream), "fino_halo_exchange")


def qkv_norm_rope_scatter(qkv: torch.Tensor, wq: Optional[torch.Tensor], wk: Optional[torch.Tensor], heads: int,
                          eps: float, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], dst_ptrs, world: int,
                          rank: int, rows_per_rank: int, dst_row_stride: int) -> None:
    """RMSNorm-across-heads (+ Wan RoPE) of the local [rows, 3*D] projections, stored into every rank's exchange buffer
    (``fino_qkv_norm_rope_scatter``). cos/sin: fp32 [rows, head_dim] rows of the local tokens."""
    assert qkv.dtype == torch.bfloat16
    lib, stream = _prep(qkv, wq, wk, cos, sin)
    rows, cols, stride = _rows2d(qkv)
    dim = cols // 3
    assert dim * 3 == cols and dim % heads == 0
    head_dim = dim // heads
    for t in (wq, wk):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == dim
    if cos is not None:
        assert cos.dtype == sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape == (rows, head_dim) and sin.shape == (rows, head_dim)
    status = lib.fino_qkv_norm_rope_scatter(qkv.data_ptr(), rows, stride, _ptr(wq), _ptr(wk), heads, head_dim,
                                            float(eps), _ptr(cos), _ptr(sin), dst_ptrs, world, rank, rows_per_rank,
                                            dst_row_stride, stream)
    _lib.check(status, "fino_qkv_norm_rope_scatter")


def qkv_ln_rope_scatter(qkv: torch.Tensor, wq, bq, wk, bk, heads: int, eps: float, cos: Optional[torch.Tensor],
                        sin: Optional[torch.Tensor], rope_skip: int, dst_ptrs, world: int, rank: int, rows_per_rank: int,
                        dst_row_stride: int) -> None:
    """CogVideoX: per-head LayerNorm(64) (+ RoPE on local rows >= rope_skip) of the local [rows, 3*D] projections,
    stored into every rank's exchange buffer (``fino_qkv_ln_rope_scatter``). cos/sin: fp32 [rows - rope_skip, 64]."""
    assert qkv.dtype == torch.bfloat16
    lib, stream = _prep(qkv, wq, bq, wk, bk, cos, sin)
    rows, cols, stride = _rows2d(qkv)
    dim = cols // 3
    assert dim * 3 == cols and dim == heads * 64
    for t in (wq, bq, wk, bk):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == 64
    if cos is not None:
        assert cos.dtype == sin.dtype == torch.float32 and cos.is_contiguous() and sin.is_contiguous()
        assert cos.shape == sin.shape == (rows - rope_skip, 64), (tuple(cos.shape), rows, rope_skip)
    status = lib.fino_qkv_ln_rope_scatter(qkv.data_ptr(), rows, stride, _ptr(wq), _ptr(bq), _ptr(wk), _ptr(bk), heads, 64,
                                          float(eps), _ptr(cos), _ptr(sin), rope_skip, dst_ptrs, world, rank,
                                          rows_per_rank, dst_row_stride, stream)
    _lib.check(status, "fino_qkv_ln_rope_scatter")


def attention_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, o_ptrs, num_owners: int,
                      rows_per_owner: int, o_row_stride: int, scale: Optional[float] = None) -> None:
    """``attention`` whose output rows are stored into their owners' buffers (``fino_attention_fwd_scatter``):
    row g -> o_ptrs[g // rows_per_owner] + (g % rows_per_owner) * o_row_stride. Batch 1."""
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    lib, stream = _prep(q, k, v)
    b, nq, inner = q.shape
    nk = k.shape[1]
    d = inner // heads
    assert b == 1 and d in (64, 128), "scatter attention: batch 1, head_dim 64 or 128"
    for t in (q, k, v):
        assert t.stride(2) == 1
    if scale is None:
        scale = d ** -0.5
    status = lib.fino_attention_fwd_scatter(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), o_ptrs, num_owners, rows_per_owner, b, heads, nq, nk, d,
        q.stride(1), k.stride(1), v.stride(1), o_row_stride, q.stride(0), k.stride(0), v.stride(0), 0, float(scale),
        stream)
    _lib.check(status, "fino_attention_fwd_scatter")
