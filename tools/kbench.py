#!/usr/bin/env python
"""Isolated kernel timings at the BASELINE config-2 / config-3 shapes (CUDA events on the launching stream, 3 warm-up +
N timed back-to-back launches per kernel). For A/B-ing a kernel change; the judged numbers come from bench.py.
Usage: python tools/kbench.py [attn] [attn64] [cross] [gemm] [rows] [--iters 10]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frameino_b200 import ops  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
iters = 10
if "--iters" in sys.argv:
    iters = int(sys.argv[sys.argv.index("--iters") + 1])
which = args or ["attn", "attn64", "cross", "gemm", "rows"]  # also: gemm8 rows64 vae
if os.environ.get("FINO_GEMM_MODE"):
    ops.gemm_set_mode(int(os.environ["FINO_GEMM_MODE"]))
if os.environ.get("FINO_ATTN_VARIANT"):
    ops.attention_set_variant(int(os.environ["FINO_ATTN_VARIANT"]))


def timeit(fn, flops=0.0, bytes_=0.0, name=""):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    extra = ""
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    print(f"KBENCH {name:<44} {ms * 1e3:10.1f} us{extra}", flush=True)


n, h, hd = 28160, 24, 128
d = h * hd
torch.manual_seed(0)
if "attn" in which:
    qkv = torch.randn(1, n, 3 * d, device="cuda").bfloat16()
    o = torch.empty(1, n, d, device="cuda", dtype=torch.bfloat16)
    timeit(lambda: ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h, out=o), 4.0 * n * n * d,
           name="attention d128 28160x28160 h24")
    n8 = n
    q8 = qkv[..., : 3 * 384].contiguous()  # what one rank sees at P = 8: 3 heads, all tokens
    timeit(lambda: ops.attention(q8[..., :384], q8[..., 384:768], q8[..., 768:], 3), 4.0 * n8 * n8 * 384,
           name="attention d128 28160x28160 h3 (P=8 shard)")
    q4 = qkv[..., : 3 * 768].contiguous()  # P = 4: 6 heads
    timeit(lambda: ops.attention(q4[..., :768], q4[..., 768:1536], q4[..., 1536:], 6), 4.0 * n8 * n8 * 768,
           name="attention d128 28160x28160 h6 (P=4 shard)")
    ops.attention_set_split(0)
    timeit(lambda: ops.attention(q8[..., :384], q8[..., 384:768], q8[..., 768:], 3), 4.0 * n8 * n8 * 384,
           name="  same h3, KV split off")
    timeit(lambda: ops.attention(q4[..., :768], q4[..., 768:1536], q4[..., 1536:], 6), 4.0 * n8 * n8 * 768,
           name="  same h6, KV split off")
    ops.attention_set_split(-1)
    del qkv, o
if "attn64" in which:
    n2, h2 = 19126, 48
    qkv = torch.randn(1, n2, 3 * d, device="cuda").bfloat16()
    timeit(lambda: ops.attention(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h2), 4.0 * n2 * n2 * d,
           name="attention d64 19126x19126 h48")
    del qkv
if "cross" in which:
    q = torch.randn(1, n, d, device="cuda").bfloat16()
    kv = torch.randn(1, 512, 2 * d, device="cuda").bfloat16()
    timeit(lambda: ops.attention(q, kv[..., :d], kv[..., d:], h), 4.0 * n * 512 * d, name="cross attention 28160x512 h24")
    del q, kv
if "gemm8" in which:  # the per-rank GEMM shapes at 8-way sequence parallel (M = 3520)
    m8 = n // 8
    f = 14336
    a = torch.randn(m8, d, device="cuda").bfloat16()
    x = a.clone()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.ones(m8, device="cuda", dtype=torch.int32)
    for (nn, kk, epi, nm) in [(3 * d, d, ops.EPI_NONE, "qkv"), (d, d, ops.EPI_GATE_RESIDUAL, "out+gate-res"),
                              (d, d, ops.EPI_NONE, "cross-q"), (f, d, ops.EPI_GELU_TANH, "ffn-up+gelu"),
                              (d, f, ops.EPI_GATE_RESIDUAL, "ffn-down+gate-res")]:
        inp = a if kk == d else torch.randn(m8, kk, device="cuda").bfloat16()
        w = (torch.randn(nn, kk, device="cuda") / math.sqrt(kk)).bfloat16()
        b = torch.randn(nn, device="cuda").bfloat16()
        if epi == ops.EPI_GATE_RESIDUAL:
            fn = lambda: ops.linear(inp, w, b, epilogue=epi, residual=x, gate=tab[:, :d], row_index=ridx, out=x)  # noqa: E731
        else:
            out = torch.empty(m8, nn, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.linear(inp, w, b, epilogue=epi, out=out)  # noqa: E731
        timeit(fn, 2.0 * m8 * nn * kk, name=f"gemm {nm} M{m8} N{nn} K{kk}")
        del w, b
    del a, x
if "gemm" in which:
    a = torch.randn(n, d, device="cuda").bfloat16()
    f = 14336
    x = a.clone()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.zeros(n, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    for (nn, kk, epi, nm) in [(3 * d, d, ops.EPI_NONE, "qkv"), (d, d, ops.EPI_GATE_RESIDUAL, "out+gate-res"),
                              (d, d, ops.EPI_NONE, "cross-q"), (f, d, ops.EPI_GELU_TANH, "ffn-up+gelu"),
                              (d, f, ops.EPI_GATE_RESIDUAL, "ffn-down+gate-res")]:
        inp = a if kk == d else torch.randn(n, kk, device="cuda").bfloat16()
        w = (torch.randn(nn, kk, device="cuda") / math.sqrt(kk)).bfloat16()
        b = torch.randn(nn, device="cuda").bfloat16()
        if epi == ops.EPI_GATE_RESIDUAL:
            fn = lambda: ops.linear(inp, w, b, epilogue=epi, residual=x, gate=tab[:, :d], row_index=ridx, out=x)  # noqa: E731
        else:
            out = torch.empty(n, nn, device="cuda", dtype=torch.bfloat16)
            fn = lambda: ops.linear(inp, w, b, epilogue=epi, out=out)  # noqa: E731
        timeit(fn, 2.0 * n * nn * kk, name=f"gemm {nm} M{n} N{nn} K{kk}")
        del w, b
    del a, x
if "rows" in which:
    x = torch.randn(n, d, device="cuda").bfloat16()
    tab = torch.randn(2, 6 * d, device="cuda")
    ridx = torch.zeros(n, device="cuda", dtype=torch.int32)
    ridx[880:] = 1
    o = torch.empty_like(x)
    g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    for variant in (0, 3):
        ops.rows_set_variant(variant, 2)
        timeit(lambda: ops.ln_modulate(x, 1e-6, shift=tab[:, :d], scale=tab[:, d:2 * d], row_index=ridx, out=o),
               bytes_=2.0 * n * d * 2, name=f"ln_modulate 28160x3072 (variant {variant})")
        timeit(lambda: ops.ln_modulate(x, 1e-6, gamma=g, beta=b, out=o), bytes_=2.0 * n * d * 2,
               name=f"ln affine 28160x3072 (variant {variant})")
        x8 = x[:3520]
        timeit(lambda: ops.ln_modulate(x8, 1e-6, shift=tab[:, :d], scale=tab[:, d:2 * d], row_index=ridx[:3520],
                                       out=o[:3520]), bytes_=2.0 * 3520 * d * 2,
               name=f"ln_modulate 3520x3072 (variant {variant})")
    qkv = torch.randn(1, n, 3 * d, device="cuda").bfloat16()
    wq = torch.ones(d, device="cuda").bfloat16()
    cos = torch.rand(n, hd, device="cuda")
    sin = torch.rand(n, hd, device="cuda")
    for variant in (1, 2):
        ops.rows_set_variant(3, variant)
        timeit(lambda: ops.qk_norm_rope(qkv[..., :d], wq, qkv[..., d:2 * d], wq, h, rope_mode=ops.ROPE_WAN, cos=cos,
                                        sin=sin, seq_len=n), bytes_=4.0 * n * d * 2 + 2.0 * n * hd * 4,
               name=f"qk_norm_rope 28160x(2x3072) (variant {variant})")
        timeit(lambda: ops.qk_norm_rope(x, wq, None, None, h), bytes_=2.0 * n * d * 2,
               name=f"q_norm (cross) 28160x3072 (variant {variant})")
if "rows64" in which:  # CogVideoX: per-head LayerNorm(64) + RoPE on the joint [226 text + video] sequence
    n2, h2, hd2 = 19126, 48, 64
    qkv = torch.randn(1, n2, 3 * d, device="cuda").bfloat16()
    w64, b64 = torch.ones(hd2, device="cuda").bfloat16(), torch.zeros(hd2, device="cuda").bfloat16()
    cos = torch.rand(n2 - 226, hd2, device="cuda")
    sin = torch.rand(n2 - 226, hd2, device="cuda")
    for variant in (0, 2, 3, 4, 5, 6):  # 0 generic warp-per-row; ln64 kernel with G groups in flight: 2 (default) G=4; 3: 12; 4: 6; 5: 3; 6: 2
        ops.rows_set_variant(3, variant)
        timeit(lambda: ops.qk_norm_rope(qkv[..., :d], w64, qkv[..., d:2 * d], w64, h2, b0=b64, b1=b64,
                                        norm_mode=ops.QK_LAYERNORM_PER_HEAD, rope_mode=ops.ROPE_COGVIDEOX, cos=cos,
                                        sin=sin, seq_len=n2, rope_skip=226),
               bytes_=4.0 * n2 * d * 2 + 2.0 * n2 * hd2 * 4, name=f"qk LayerNorm(64)+RoPE 19126x(2x3072) (variant {variant})")
    ops.rows_set_variant(3, 2)
print("done", which)
if "vae" in which:  # Wan VAE row kernels at the decoder's widest stages (4 frames of 352x640x256, 176x320x512, ...)
    for rows_, c_ in [(4 * 352 * 640, 256), (4 * 176 * 320, 512), (2 * 88 * 160, 1024), (4 * 352 * 640, 160)]:
        xx = torch.randn(rows_, c_, device="cuda").bfloat16()
        oo = torch.empty_like(xx)
        gm = torch.ones(c_, device="cuda")
        timeit(lambda: ops.rms_act_cl(xx, gm, silu=True, out=oo), bytes_=2.0 * rows_ * c_ * 2, name=f"rms_act_cl {rows_}x{c_}")
        del xx, oo
    yy = torch.randn(4, 352, 640, 512, device="cuda").bfloat16()
    ss = torch.randn(4, 176, 320, 1024, device="cuda").bfloat16()
    timeit(lambda: ops.dupup_add_cl(yy, ss, 1, 2, False), bytes_=2.0 * yy.numel() * 2 + ss.numel() * 2, name="dupup_add_cl 4x352x640x512")
    timeit(lambda: ops.upsample2x_cl(ss), bytes_=5.0 * ss.numel() * 2, name="upsample2x_cl 4x176x320x1024")
    del yy, ss
