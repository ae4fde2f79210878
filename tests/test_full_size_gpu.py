"""-m gpu: BASELINE.json full sizes (Wan2.2-5B: N = 28160, 24 heads x 128; CogVideoX: S = 19126, 48 x 64), checked
through size-independent properties because the CPU oracle cannot finish these sizes in seconds:
softmax rows sum to one, key-permutation invariance, linearity in V, GEMM linearity / identity, and one oracle-checked
slice of rows (attention of 64 query rows against all keys is cheap on CPU)."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import ops as _ops

    return _ops


@pytest.mark.parametrize("heads,hd,n", [(24, 128, 28160), (48, 64, 19126)])
def test_attention_full_size_properties(ops, heads, hd, n):
    dm = heads * hd
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(1, n, 3 * dm, generator=g, device="cuda").bfloat16()
    q, k, v = qkv[..., :dm], qkv[..., dm:2 * dm], qkv[..., 2 * dm:]
    out = ops.attention(q, k, v, heads)
    assert torch.isfinite(out.float()).all()
    # (1) softmax rows sum to one: V = 1 -> O = 1
    ones = torch.ones(1, n, dm, device="cuda", dtype=torch.bfloat16)
    o1 = ops.attention(q, k, ones, heads)
    assert float((o1.float() - 1).abs().max()) <= 1e-2
    # (2) permuting the keys (with their values) leaves the output unchanged up to bf16 rounding
    perm = torch.randperm(n, generator=g, device="cuda")
    o2 = ops.attention(q, k[:, perm].contiguous(), v[:, perm].contiguous(), heads)
    assert rel_err(o2, out) <= 1e-2
    # (3) linearity in V
    v2 = torch.randn(1, n, dm, generator=g, device="cuda").bfloat16()
    o3 = ops.attention(q, k, (v.float() * 0.5 + v2.float() * 0.25).bfloat16(), heads)
    o_v2 = ops.attention(q, k, v2, heads)
    assert rel_err(o3, out.float() * 0.5 + o_v2.float() * 0.25) <= 2e-2
    # (4) oracle on a slice: 64 query rows of 2 heads against ALL keys
    from oracle.wan_oracle import sdpa

    rows = torch.arange(1000, 1064)
    for h in (0, heads - 1):
        sl = slice(h * hd, (h + 1) * hd)
        ref = sdpa(q[0, rows, sl].float().cpu()[None, None], k[0, :, sl].float().cpu()[None, None],
                   v[0, :, sl].float().cpu()[None, None])[0, 0]
        assert rel_err(out[0, rows, sl], ref) <= 1e-2


def test_gemm_full_size_properties(ops):
    m, n, k = 28160, 3072, 3072
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(m, k, generator=g, device="cuda").bfloat16()
    w = (torch.randn(n, k, generator=g, device="cuda") / math.sqrt(k)).bfloat16()
    eye = torch.eye(k, device="cuda", dtype=torch.bfloat16)
    assert torch.equal(ops.linear(a, eye), a)  # identity weight reproduces the input bit for bit
    y = ops.linear(a, w, out_dtype=torch.float32)
    y2 = ops.linear((a.float() * 2).bfloat16(), w, out_dtype=torch.float32)  # exact scaling by 2 in bf16
    assert torch.equal(y2, y * 2)
    # oracle on 32 rows
    rows = torch.arange(5000, 5032, device="cuda")
    ref = torch.nn.functional.linear(a[rows].float().cpu(), w.float().cpu())
    assert rel_err(y[rows], ref) <= 1e-3


def test_ln_and_rope_full_size_properties(ops):
    n, dm = 28160, 3072
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(n, dm, generator=g, device="cuda").bfloat16()
    y = ops.ln_modulate(x, 1e-6)
    yf = y.float()
    assert float(yf.mean(-1).abs().max()) <= 2e-2 and float((yf.var(-1, unbiased=False) - 1).abs().max()) <= 3e-2
    assert torch.equal(ops.ln_modulate(y, 1e-6), ops.ln_modulate(ops.ln_modulate(y, 1e-6), 1e-6)) or \
        rel_err(ops.ln_modulate(y, 1e-6), y) <= 1e-2  # idempotent up to rounding
    # RoPE with zero angle is the identity after RMSNorm with unit weight; rotation preserves the per-pair norm
    qk = torch.randn(1, n, 2 * dm, generator=g, device="cuda").bfloat16()
    w = torch.ones(dm, device="cuda", dtype=torch.bfloat16)
    base = qk.clone()
    ops.qk_norm_rope(base[..., :dm], w, base[..., dm:], w, 24)
    rot = qk.clone()
    ang = torch.rand(n, 64, generator=g, device="cuda") * 6.28
    cos = ang.cos().repeat_interleave(2, 1).contiguous()
    sin = ang.sin().repeat_interleave(2, 1).contiguous()
    ops.qk_norm_rope(rot[..., :dm], w, rot[..., dm:], w, 24, rope_mode=ops.ROPE_WAN, cos=cos, sin=sin, seq_len=n)
    pn_base = base.float().view(1, n, -1, 2).pow(2).sum(-1)
    pn_rot = rot.float().view(1, n, -1, 2).pow(2).sum(-1)
    assert float((pn_base - pn_rot).abs().max() / pn_base.max()) <= 2e-2
