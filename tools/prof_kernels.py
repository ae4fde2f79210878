#!/usr/bin/env python
"""Runs each hot kernel a few times at the Wan2.2-5B config-2 shapes so that `ncu -k regex:<name>` can capture it.
Usage: python tools/prof_kernels.py [attn] [attn64] [gemm] [ffn] [ln] [qk]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import ops  # noqa: E402

which = sys.argv[1:] or ["attn", "gemm", "ln", "qk"]
if os.environ.get("FINO_GEMM_MODE"):
    ops.gemm_set_mode(int(os.environ["FINO_GEMM_MODE"]))
if os.environ.get("FINO_ATTN_VARIANT"):
    ops.attention_set_variant(int(os.environ["FINO_ATTN_VARIANT"]))
n, h, hd = 28160, 24, 128
d = h * hd
torch.manual_seed(0)
if "attn" in which:
    qkv = torch.randn(1, n, 3 * d, device="cuda").bfloat16()
    o = torch.empty(1, n, d, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h, out=o)
if "attn64" in which:
    n2, h2 = 19126, 48
    qkv = torch.randn(1, n2, 3 * d, device="cuda").bfloat16()
    for _ in range(3):
        ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h2)
if "gemm" in which:
    a = torch.randn(n, d, device="cuda").bfloat16()
    w = (torch.randn(3 * d, d, device="cuda") / math.sqrt(d)).bfloat16()
    b = torch.randn(3 * d, device="cuda").bfloat16()
    o = torch.empty(n, 3 * d, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(a, w, b, out=o)
if "gemm_gate" in which:  # out-projection with the gate*y + residual epilogue, in place
    a = torch.randn(n, d, device="cuda").bfloat16()
    x = torch.randn(n, d, device="cuda").bfloat16()
    w = (torch.randn(d, d, device="cuda") / math.sqrt(d)).bfloat16()
    b = torch.randn(d, device="cuda").bfloat16()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.zeros(n, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    for _ in range(3):
        ops.linear(a, w, b, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=tab[:, :d], row_index=ridx, out=x)
if "ffn" in which:
    f = 14336
    a = torch.randn(n, d, device="cuda").bfloat16()
    w = (torch.randn(f, d, device="cuda") / math.sqrt(d)).bfloat16()
    b = torch.randn(f, device="cuda").bfloat16()
    for _ in range(3):
        y = ops.linear(a, w, b, epilogue=ops.EPI_GELU_TANH)
if "ln" in which:
    x = torch.randn(n, d, device="cuda").bfloat16()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.zeros(n, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    o = torch.empty_like(x)
    for _ in range(3):
        ops.ln_modulate(x, 1e-6, shift=tab[:, :d], scale=tab[:, d:2 * d], row_index=ridx, out=o)
if "qk" in which:
    qkv = torch.randn(1, n, 3 * d, device="cuda").bfloat16()
    wq = torch.ones(d, device="cuda").bfloat16()
    cos = torch.rand(n, hd, device="cuda")
    sin = torch.rand(n, hd, device="cuda")
    for _ in range(3):
        ops.qk_norm_rope(qkv[..., :d], wq, qkv[..., d:2 * d], wq, h, rope_mode=ops.ROPE_WAN, cos=cos, sin=sin, seq_len=n)
if "conv" in which:  # Wan VAE decoder, last stage: 3x3x3 causal conv 256 -> 256 on 4 frames of 352 x 640 (+ 2 history)
    x = torch.randn(6, 352, 640, 256, device="cuda").bfloat16()
    w = (torch.randn(256, 27 * 256, device="cuda") / math.sqrt(27 * 256)).bfloat16()
    b = torch.randn(256, device="cuda").bfloat16()
    for _ in range(3):
        y = ops.conv3d_cl(x, w, b, (3, 3, 3), pad_hw=(1, 1))
    xg = torch.randn(4 * 352 * 640, 256, device="cuda").bfloat16()
    gm = torch.ones(256, device="cuda")
    for _ in range(3):
        ops.rms_act_cl(xg, gm, silu=True, out=torch.empty_like(xg))
if "ln64" in which:  # CogVideoX per-head LayerNorm(64) + RoPE
    n2, h2 = 19126, 48
    qkv = torch.randn(1, n2, 3 * d, device="cuda").bfloat16()
    w64, b64 = torch.ones(64, device="cuda").bfloat16(), torch.zeros(64, device="cuda").bfloat16()
    cos = torch.rand(n2 - 226, 64, device="cuda")
    sin = torch.rand(n2 - 226, 64, device="cuda")
    for _ in range(3):
        ops.qk_norm_rope(qkv[..., :d], w64, qkv[..., d:2 * d], w64, h2, b0=b64, b1=b64, norm_mode=ops.QK_LAYERNORM_PER_HEAD,
                         rope_mode=ops.ROPE_COGVIDEOX, cos=cos, sin=sin, seq_len=n2, rope_skip=226)
if "gemm5" in which:
    # the five GEMM shapes of a Wan block at N = 28160, this library and cuBLAS (torch) back to back in ONE session;
    # only the launches between profiler.start() / stop() are captured (ncu --profile-from-start off)
    f = 14336
    a = torch.randn(n, d, device="cuda").bfloat16()
    af = torch.randn(n, f, device="cuda").bfloat16()
    x = a.clone()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.zeros(n, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    shapes = [(3 * d, d, ops.EPI_NONE, "qkv"), (d, d, ops.EPI_GATE_RESIDUAL, "out+gate-res"), (d, d, ops.EPI_NONE, "cross-q"),
              (f, d, ops.EPI_GELU_TANH, "ffn-up+gelu"), (d, f, ops.EPI_GATE_RESIDUAL, "ffn-down+gate-res")]
    ws = [((torch.randn(nn, kk, device="cuda") / math.sqrt(kk)).bfloat16(), torch.randn(nn, device="cuda").bfloat16())
          for nn, kk, _, _ in shapes]

    def ours(i):
        nn, kk, epi, _ = shapes[i]
        inp = a if kk == d else af
        if epi == ops.EPI_GATE_RESIDUAL:
            ops.linear(inp, ws[i][0], ws[i][1], epilogue=epi, residual=x, gate=tab[:, :d], row_index=ridx, out=x)
        else:
            ops.linear(inp, ws[i][0], ws[i][1], epilogue=epi)

    def cublas(i):
        nn, kk, _, _ = shapes[i]
        torch.nn.functional.linear(a if kk == d else af, ws[i][0], ws[i][1])

    for i in range(5):
        ours(i)
        cublas(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(5):
        ours(i)
        cublas(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("done", which)
