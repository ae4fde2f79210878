#!/usr/bin/env python
"""Development harness: runs every kernel of the C-ABI library against plain torch math on the GPU, each check in its
own subprocess (a trapped kernel poisons the CUDA context) with a timeout, and writes gpurun_out/gpu_check.json.

Usage (on a GPU box):  python tools/gpu_check.py [--only NAME ...] [--timeout 180]
This is a debugging aid, not the parity suite (tests/ compares against oracle/).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rel(a, b):
    import torch

    a = a.float()
    b = b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _time_ms(fn, iters=5, warmup=2):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


# ----------------------------------------------------------------------------------------------------------------
def check_ln():
    import torch
    from frameino_b200 import ops

    torch.manual_seed(0)
    out = {}
    for rows, dim in [(37, 256), (1000, 3072), (64, 1024)]:
        x = (torch.randn(rows, dim, device="cuda") * 2 + 0.3).bfloat16()
        r = 3
        tab = torch.randn(r, 6, dim, device="cuda") * 0.5
        ridx = torch.randint(0, r, (rows,), device="cuda", dtype=torch.int32)
        shift = tab[:, 0].reshape(r, dim)
        scale = tab[:, 1].reshape(r, dim)
        y = ops.ln_modulate(x, 1e-6, shift=tab.view(r, 6 * dim)[:, 0:dim], scale=tab.view(r, 6 * dim)[:, dim:2 * dim],
                            row_index=ridx)
        ref = torch.nn.functional.layer_norm(x.float(), (dim,), eps=1e-6)
        ref = (ref * (1 + scale[ridx.long()]) + shift[ridx.long()]).bfloat16()
        out[f"mod_{rows}x{dim}"] = _rel(y, ref)
        g = torch.randn(dim, device="cuda")
        b = torch.randn(dim, device="cuda")
        y2 = ops.ln_modulate(x, 1e-6, gamma=g, beta=b)
        ref2 = torch.nn.functional.layer_norm(x.float(), (dim,), g, b, eps=1e-6).bfloat16()
        out[f"affine_{rows}x{dim}"] = _rel(y2, ref2)
        y3 = ops.ln_modulate(x, 1e-5, gamma=g, beta=b, shift=shift, scale=scale, rows_per_group=(rows + r - 1) // r,
                             bf16_steps=True)
        gi = (torch.arange(rows, device="cuda") // ((rows + r - 1) // r)).long()
        ln = torch.nn.functional.layer_norm(x, (dim,), g.bfloat16(), b.bfloat16(), eps=1e-5)
        ref3 = ln * (1 + scale.bfloat16()[gi]) + shift.bfloat16()[gi]
        out[f"bf16steps_{rows}x{dim}"] = _rel(y3, ref3)
    x = torch.randn(28160, 3072, device="cuda").bfloat16()
    tab = torch.randn(2, 6 * 3072, device="cuda")
    ridx = torch.zeros(28160, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    o = torch.empty_like(x)
    ms = _time_ms(lambda: ops.ln_modulate(x, 1e-6, shift=tab[:, :3072], scale=tab[:, 3072:6144], row_index=ridx, out=o))
    out["time_ms_28160x3072"] = ms
    out["gbps"] = 2 * x.numel() * 2 / ms / 1e6
    return out


def check_gate():
    import torch
    from frameino_b200 import ops

    torch.manual_seed(1)
    rows, dim, r = 515, 768, 2
    x = torch.randn(rows, dim, device="cuda").bfloat16()
    y = torch.randn(rows, dim, device="cuda").bfloat16()
    gate = torch.randn(r, dim, device="cuda")
    ridx = torch.randint(0, r, (rows,), device="cuda", dtype=torch.int32)
    o = ops.gate_residual(x, y, gate, row_index=ridx)
    ref = (x.float() + y * gate[ridx.long()]).bfloat16()
    o2 = ops.gate_residual(x, y)
    return {"gated": _rel(o, ref), "plain": _rel(o2, x + y)}


def _wan_rope_ref(x, cos, sin):
    import torch

    xr = x.view(*x.shape[:-1], -1, 2)
    x1, x2 = xr[..., 0], xr[..., 1]
    c = cos[..., 0::2]
    s = sin[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x1 * c - x2 * s
    out[..., 1::2] = x1 * s + x2 * c
    return out.type_as(x)


def check_qk():
    import torch
    from frameino_b200 import ops

    torch.manual_seed(2)
    res = {}
    # Wan: RMS across heads + rope, q/k are column slices of a fused [rows, 3D] buffer
    for (b, n, h, d) in [(1, 384, 8, 32), (2, 300, 24, 128)]:
        D = h * d
        qkv = torch.randn(b, n, 3 * D, device="cuda").bfloat16()
        wq = (1 + 0.1 * torch.randn(D, device="cuda")).bfloat16()
        wk = (1 + 0.1 * torch.randn(D, device="cuda")).bfloat16()
        ang = torch.rand(n, d // 2, device="cuda") * 6.28
        cos = ang.cos().repeat_interleave(2, dim=1).contiguous()
        sin = ang.sin().repeat_interleave(2, dim=1).contiguous()
        q0 = qkv[..., :D].clone()
        k0 = qkv[..., D:2 * D].clone()

        def rms(x, w):
            v = x.float().pow(2).mean(-1, keepdim=True)
            return (x.float() * torch.rsqrt(v + 1e-6)).to(w.dtype) * w

        def ref(x, w):
            y = rms(x, w).view(b, n, h, d).transpose(1, 2)
            y = _wan_rope_ref(y, cos[None, None], sin[None, None])
            return y.transpose(1, 2).reshape(b, n, D)

        rq, rk = ref(q0, wq), ref(k0, wk)
        ops.qk_norm_rope(qkv[..., :D], wq, qkv[..., D:2 * D], wk, h, rope_mode=ops.ROPE_WAN, cos=cos, sin=sin, seq_len=n)
        res[f"wan_q_{b}x{n}x{h}x{d}"] = _rel(qkv[..., :D], rq)
        res[f"wan_k_{b}x{n}x{h}x{d}"] = _rel(qkv[..., D:2 * D], rk)
    # CogVideoX: per-head LayerNorm + rope on tokens >= text_len
    b, s, h, d, text = 2, 226 + 150, 48, 64, 226
    D = h * d
    qkv = torch.randn(b, s, 3 * D, device="cuda").bfloat16()
    wq, bq = (1 + 0.1 * torch.randn(d, device="cuda")).bfloat16(), (0.1 * torch.randn(d, device="cuda")).bfloat16()
    wk, bk = (1 + 0.1 * torch.randn(d, device="cuda")).bfloat16(), (0.1 * torch.randn(d, device="cuda")).bfloat16()
    ang = torch.rand(s - text, d // 2, device="cuda") * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).contiguous()
    sin = ang.sin().repeat_interleave(2, dim=1).contiguous()

    def cog_ref(x, w, bb):
        y = x.view(b, s, h, d).transpose(1, 2)
        y = torch.nn.functional.layer_norm(y, (d,), w, bb, eps=1e-6)
        xi = y[:, :, text:]
        xr, xim = xi.reshape(*xi.shape[:-1], -1, 2).unbind(-1)
        rot = torch.stack([-xim, xr], dim=-1).flatten(3)
        o = (xi.float() * cos[None, None] + rot.float() * sin[None, None]).to(xi.dtype)
        y = y.clone()
        y[:, :, text:] = o
        return y.transpose(1, 2).reshape(b, s, D)

    rq = cog_ref(qkv[..., :D].clone(), wq, bq)
    rk = cog_ref(qkv[..., D:2 * D].clone(), wk, bk)
    ops.qk_norm_rope(qkv[..., :D], wq, qkv[..., D:2 * D], wk, h, b0=bq, b1=bk, norm_mode=ops.QK_LAYERNORM_PER_HEAD,
                     rope_mode=ops.ROPE_COGVIDEOX, cos=cos, sin=sin, seq_len=s, rope_skip=text)
    res["cog_q"] = _rel(qkv[..., :D], rq)
    res["cog_k"] = _rel(qkv[..., D:2 * D], rk)
    # timing at Wan config-2 size
    n, h, d = 28160, 24, 128
    D = h * d
    qkv = torch.randn(1, n, 3 * D, device="cuda").bfloat16()
    wq = torch.ones(D, device="cuda").bfloat16()
    cos = torch.rand(n, d, device="cuda")
    sin = torch.rand(n, d, device="cuda")
    ms = _time_ms(lambda: ops.qk_norm_rope(qkv[..., :D], wq, qkv[..., D:2 * D], wq, h, rope_mode=ops.ROPE_WAN, cos=cos,
                                           sin=sin, seq_len=n))
    res["time_ms_wan_28160"] = ms
    res["gbps"] = 4 * n * D * 2 / ms / 1e6
    return res


def check_misc():
    import torch
    from frameino_b200 import ops

    torch.manual_seed(3)
    res = {}
    # patchify == conv3d
    b, c, f, h, w, dm = 1, 32, 6, 16, 16, 256
    x = torch.randn(b, c, f, h, w, device="cuda").bfloat16()
    conv = torch.nn.Conv3d(c, dm, (1, 2, 2), (1, 2, 2)).cuda().bfloat16()
    rows = ops.patchify(x, (b, c, f, h, w), x.stride(), (1, 2, 2))
    y = ops.linear(rows, conv.weight.view(dm, -1), conv.bias)
    ref = conv(x).flatten(2).transpose(1, 2).reshape(-1, dm)
    res["patchify_conv3d"] = _rel(y, ref)
    # cog layout [B,F,C,H,W], conv2d per frame
    xc = torch.randn(2, 3, 16, 8, 12, device="cuda").bfloat16()
    conv2 = torch.nn.Conv2d(16, 128, 2, 2).cuda().bfloat16()
    st = xc.stride()
    rows = ops.patchify(xc, (2, 16, 3, 8, 12), (st[0], st[2], st[1], st[3], st[4]), (1, 2, 2))
    y = ops.linear(rows, conv2.weight.view(128, -1), conv2.bias)
    ref = conv2(xc.reshape(-1, 16, 8, 12)).view(2, 3, 128, -1).transpose(2, 3).reshape(-1, 128)
    res["patchify_conv2d"] = _rel(y, ref)
    # unpatchify wan
    B, F, H, W, C = 1, 6, 16, 16, 16
    r = torch.randn(B * F * (H // 2) * (W // 2), 4 * C, device="cuda").bfloat16()
    ref = r.reshape(B, F, H // 2, W // 2, 1, 2, 2, -1).permute(0, 7, 1, 4, 2, 5, 3, 6).flatten(6, 7).flatten(4, 5).flatten(2, 3)
    out = torch.empty(B, C, F, H, W, device="cuda", dtype=torch.bfloat16)
    ops.unpatchify(r, out, (B, C, F, H, W), out.stride(), (1, 2, 2), True)
    res["unpatchify_wan"] = float((out.float() - ref.float()).abs().max())
    # unpatchify cog: rows [B,F,h/2,w/2, C*2*2] -> [B,F,C,H,W]
    r = torch.randn(2 * 3 * 4 * 6, 16 * 4, device="cuda").bfloat16()
    ref = r.reshape(2, 3, 4, 6, -1, 2, 2).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    out = torch.empty(2, 3, 16, 8, 12, device="cuda", dtype=torch.bfloat16)
    st = out.stride()
    ops.unpatchify(r, out, (2, 16, 3, 8, 12), (st[0], st[2], st[1], st[3], st[4]), (1, 2, 2), False)
    res["unpatchify_cog"] = float((out.float() - ref.float()).abs().max())
    # timestep embedding
    t = torch.tensor([0.0, 500.0, 999.0, 37.5], device="cuda")
    e = ops.timestep_embedding(t, 256)
    half = 128
    ex = torch.exp(-math.log(10000) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    a = t[:, None] * ex[None]
    ref = torch.cat([a.cos(), a.sin()], -1)
    res["timestep_embedding"] = float((e - ref).abs().max())
    # small-M linear
    xs = torch.randn(2, 256, device="cuda")
    wf = torch.randn(3072, 256, device="cuda") * 0.05
    bf = torch.randn(3072, device="cuda") * 0.1
    y = ops.linear_small_m(xs, wf, bf, act_out=1)
    ref = torch.nn.functional.silu(xs @ wf.t() + bf)
    res["small_m_f32"] = _rel(y, ref)
    wb = wf.bfloat16()
    bb = bf.bfloat16()
    y = ops.linear_small_m(xs, wb, bb, act_in=1, round_in=True, round_out=True)
    ref = (torch.nn.functional.silu(xs).bfloat16() @ wb.t() + bb)
    res["small_m_bf16"] = _rel(y, ref)
    # mod table
    tab = torch.randn(5, 6 * 64, device="cuda")
    proj = torch.randn(3, 6 * 64, device="cuda")
    o = ops.build_mod_table(tab, proj, 5, 6 * 64)
    res["mod_table"] = float((o - (tab[:, None] + proj[None])).abs().max())
    return res


def check_gemm():
    import torch
    from frameino_b200 import ops

    torch.manual_seed(4)
    res = {}
    for (m, n, k) in [(128, 256, 64), (128, 128, 128), (256, 512, 256), (300, 192, 384), (1000, 3072, 3072), (2, 3072, 256)]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
        bias = torch.randn(n, device="cuda").bfloat16()
        y = ops.linear(a, w, bias)
        ref = (a.float() @ w.float().t() + bias.float())
        res[f"plain_{m}x{n}x{k}"] = _rel(y, ref)
    m, n, k = 777, 1024, 512
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
    bias = torch.randn(n, device="cuda").bfloat16()
    lin = (a.float() @ w.float().t() + bias.float()).bfloat16()
    res["gelu"] = _rel(ops.linear(a, w, bias, epilogue=ops.EPI_GELU_TANH),
                       torch.nn.functional.gelu(lin, approximate="tanh"))
    res["silu"] = _rel(ops.linear(a, w, bias, epilogue=ops.EPI_SILU), torch.nn.functional.silu(lin))
    res["fp32_out"] = _rel(ops.linear(a, w, bias, out_dtype=torch.float32), a.float() @ w.float().t() + bias.float())
    x = torch.randn(m, n, device="cuda").bfloat16()
    gate = torch.randn(3, n, device="cuda")
    ridx = torch.randint(0, 3, (m,), device="cuda", dtype=torch.int32)
    y = ops.linear(a, w, bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=gate, row_index=ridx)
    res["gate_residual"] = _rel(y, (x.float() + lin * gate[ridx.long()]).bfloat16())
    y = ops.linear(a, w, bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=x)
    res["residual_only"] = _rel(y, x + lin)
    # in-place residual (out aliases residual), as the block uses it
    xc = x.clone()
    ops.linear(a, w, bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=xc, gate=gate, row_index=ridx, out=xc)
    res["gate_residual_inplace"] = _rel(xc, (x.float() + lin * gate[ridx.long()]).bfloat16())
    # performance at Wan config-2 shapes
    for (m, n, k) in [(28160, 3072, 3072), (28160, 9216, 3072), (28160, 14336, 3072), (28160, 3072, 14336)]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
        bias = torch.randn(n, device="cuda").bfloat16()
        o = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        ms = _time_ms(lambda: ops.linear(a, w, bias, out=o))
        res[f"tflops_{m}x{n}x{k}"] = 2.0 * m * n * k / ms / 1e9
        ms_t = _time_ms(lambda: torch.nn.functional.linear(a, w, bias))
        res[f"tflops_torch_{m}x{n}x{k}"] = 2.0 * m * n * k / ms_t / 1e9
        if n == 3072 and k == 3072:
            res[f"rel_big_{m}x{n}x{k}"] = _rel(o, torch.nn.functional.linear(a, w, bias))
    return res


def _attn_ref(q, k, v, heads, scale=None):
    import torch

    b, nq, inner = q.shape
    d = inner // heads
    qh = q.view(b, nq, heads, d).transpose(1, 2).float()
    kh = k.view(b, -1, heads, d).transpose(1, 2).float()
    vh = v.view(b, -1, heads, d).transpose(1, 2).float()
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    return o.transpose(1, 2).reshape(b, nq, inner)


def check_attn(hd=128):
    import torch
    from frameino_b200 import ops

    torch.manual_seed(5)
    res = {}
    cases = [(1, 1, 256, 128), (1, 1, 256, 256), (1, 2, 512, 1024), (2, 3, 1000, 1000), (1, 2, 300, 77), (1, 4, 4096, 4096)]
    for (b, h, nq, nk) in cases:
        D = h * hd
        q = torch.randn(b, nq, D, device="cuda").bfloat16()
        k = torch.randn(b, nk, D, device="cuda").bfloat16()
        v = torch.randn(b, nk, D, device="cuda").bfloat16()
        o = ops.attention(q, k, v, h)
        ref = _attn_ref(q, k, v, h)
        res[f"d{hd}_b{b}h{h}_{nq}x{nk}"] = _rel(o, ref)
    # strided: q/k/v as column blocks of a fused qkv buffer; larger-magnitude scores to exercise the lazy rescale
    b, h, n = 1, 3, 1536
    D = h * hd
    qkv = (torch.randn(b, n, 3 * D, device="cuda") * 2.5).bfloat16()
    o = ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], h)
    ref = _attn_ref(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], h)
    res[f"d{hd}_fused_qkv_peaky"] = _rel(o, ref)
    return res


def check_attn128():
    return check_attn(128)


def check_attn64():
    return check_attn(64)


def check_attn_perf():
    import torch
    from frameino_b200 import ops

    res = {}
    for (h, hd, n) in [(24, 128, 28160), (48, 64, 19126)]:
        D = h * hd
        qkv = torch.randn(1, n, 3 * D, device="cuda").bfloat16()
        o = torch.empty(1, n, D, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], h, out=o)
        ms = _time_ms(f, iters=3, warmup=1)
        res[f"ms_h{h}d{hd}_n{n}"] = ms
        res[f"tflops_h{h}d{hd}_n{n}"] = 4.0 * n * n * D / ms / 1e9
        qh = qkv[..., :D].view(1, n, h, hd).transpose(1, 2)
        kh = qkv[..., D:2 * D].view(1, n, h, hd).transpose(1, 2)
        vh = qkv[..., 2 * D:].view(1, n, h, hd).transpose(1, 2)
        g = lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
        ms_t = _time_ms(g, iters=3, warmup=1)
        res[f"tflops_torch_sdpa_h{h}d{hd}_n{n}"] = 4.0 * n * n * D / ms_t / 1e9
        ref = g().transpose(1, 2).reshape(1, n, D)
        res[f"rel_vs_sdpa_h{h}d{hd}"] = _rel(o, ref)
    return res


def check_gemm2():
    """CTA-pair (cta_group::2) GEMM vs the single-CTA kernel: correctness on ragged shapes + throughput."""
    import torch
    from frameino_b200 import ops

    torch.manual_seed(6)
    res = {}
    for (m, n, k) in [(256, 256, 64), (512, 512, 256), (300, 264, 384), (1000, 3072, 3072), (5000, 192, 512),
                      (2, 3072, 256), (28160, 3072, 3072)]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
        bias = torch.randn(n, device="cuda").bfloat16()
        ref = torch.nn.functional.linear(a, w, bias).float()
        for mode in (1, 2, 3):
            ops.gemm_set_mode(mode)
            y = ops.linear(a, w, bias)
            torch.cuda.synchronize()
            res[f"mode{mode}_{m}x{n}x{k}"] = _rel(y, ref)
    m, n, k = 777, 1024, 512
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
    bias = torch.randn(n, device="cuda").bfloat16()
    x = torch.randn(m, n, device="cuda").bfloat16()
    gate = torch.randn(3, n, device="cuda")
    ridx = torch.randint(0, 3, (m,), device="cuda", dtype=torch.int32)
    lin = (a.float() @ w.float().t() + bias.float()).bfloat16()
    ops.gemm_set_mode(2)
    y = ops.linear(a, w, bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=gate, row_index=ridx)
    res["mode2_gate_residual"] = _rel(y, (x.float() + lin * gate[ridx.long()]).bfloat16())
    res["mode2_gelu"] = _rel(ops.linear(a, w, bias, epilogue=ops.EPI_GELU_TANH),
                             torch.nn.functional.gelu(lin, approximate="tanh"))
    for (m, n, k) in [(28160, 3072, 3072), (28160, 9216, 3072), (28160, 14336, 3072), (28160, 3072, 14336)]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") / math.sqrt(k)).bfloat16()
        bias = torch.randn(n, device="cuda").bfloat16()
        o = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        for mode in (1, 2, 3):
            ops.gemm_set_mode(mode)
            ms = _time_ms(lambda: ops.linear(a, w, bias, out=o), iters=10, warmup=3)
            res[f"tflops_mode{mode}_{m}x{n}x{k}"] = 2.0 * m * n * k / ms / 1e9
        ms_t = _time_ms(lambda: torch.nn.functional.linear(a, w, bias), iters=10, warmup=3)
        res[f"tflops_torch_{m}x{n}x{k}"] = 2.0 * m * n * k / ms_t / 1e9
    ops.gemm_set_mode(0)
    return res


def check_attn_variants():
    """Scheduling variants of the attention kernel: accuracy vs fp32 SDPA and throughput at the config-2/3 shapes."""
    import torch
    from frameino_b200 import ops

    torch.manual_seed(7)
    res = {}
    for hd, h, n in [(128, 24, 28160), (64, 48, 19126)]:
        D = h * hd
        qkv = torch.randn(1, n, 3 * D, device="cuda").bfloat16()
        o = torch.empty(1, n, D, device="cuda", dtype=torch.bfloat16)
        # accuracy on a smaller, peaky problem
        b2, h2, n2 = 1, 3, 1500
        small = (torch.randn(b2, n2, 3 * h2 * hd, device="cuda") * 2.0).bfloat16()
        D2 = h2 * hd
        ref = _attn_ref(small[..., :D2], small[..., D2:2 * D2], small[..., 2 * D2:], h2)
        for v in (1, 2, 3, 4, 5):
            ops.attention_set_variant(v)
            out = ops.attention(small[..., :D2], small[..., D2:2 * D2], small[..., 2 * D2:], h2)
            res[f"d{hd}_v{v}_rel"] = _rel(out, ref)
            f = lambda: ops.attention(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], h, out=o)
            ms = _time_ms(f, iters=4, warmup=1)
            res[f"d{hd}_v{v}_tflops"] = 4.0 * n * n * D / ms / 1e9
    ops.attention_set_variant(0)
    return res


CHECKS = {
    "attn_variants": check_attn_variants,
    "gemm2": check_gemm2,
    "ln": check_ln,
    "gate": check_gate,
    "qk": check_qk,
    "misc": check_misc,
    "gemm": check_gemm,
    "attn128": check_attn128,
    "attn64": check_attn64,
    "attn_perf": check_attn_perf,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--child", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_check.json"))
    args = ap.parse_args()
    if args.child:
        res = CHECKS[args.child]()
        print("RESULT " + json.dumps(res))
        return
    names = args.only or list(CHECKS)
    report = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], capture_output=True,
                               text=True, timeout=args.timeout, cwd=ROOT)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if p.returncode == 0 and line:
                report[name] = json.loads(line[-1][7:])
            else:
                report[name] = {"error": f"rc={p.returncode}", "stdout": p.stdout[-3000:], "stderr": p.stderr[-3000:]}
        except subprocess.TimeoutExpired as e:
            report[name] = {"error": "timeout", "stdout": (e.stdout or b"")[-3000:].decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-3000:]}
        report[name + "_seconds"] = round(time.time() - t0, 1)
        print(f"== {name}: {json.dumps(report[name], indent=1)[:6000]}", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
