"""-m gpu: every C-ABI kernel against the CPU oracle's primitives on the same seeded inputs (bf16 data; the oracle
computes in fp32 / with the reference's cast points). Tolerances are bf16 rounding: rel err = max|a-b| / max|b|."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2  # one bf16 ulp is 2^-8 = 3.9e-3 relative; results are compared after an independent rounding


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import ops as _ops

    return _ops


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).bfloat16()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,dim", [(1, 8), (37, 256), (513, 3072), (64, 5120)])
def test_ln_modulate_vs_oracle(ops, rows, dim):
    from oracle.wan_oracle import fp32_layer_norm

    x = bf(rows, dim, scale=2.0)
    r = 3
    tab = torch.randn(r, 6 * dim, generator=torch.Generator().manual_seed(1)) * 0.5
    ridx = torch.randint(0, r, (rows,), generator=torch.Generator().manual_seed(2)).int()
    shift, scale = tab[:, :dim], tab[:, dim:2 * dim]
    ref = (fp32_layer_norm(x.float(), None, None, 1e-6) * (1 + scale[ridx.long()]) + shift[ridx.long()]).bfloat16()
    tc = tab.cuda()
    out = ops.ln_modulate(x.cuda(), 1e-6, shift=tc[:, :dim], scale=tc[:, dim:2 * dim], row_index=ridx.cuda())
    assert rel_err(out, ref) <= BF16_TOL
    g, b = torch.randn(dim), torch.randn(dim)
    ref2 = fp32_layer_norm(x.float(), g, b, 1e-6).bfloat16()
    out2 = ops.ln_modulate(x.cuda(), 1e-6, gamma=g.cuda(), beta=b.cuda())
    assert rel_err(out2, ref2) <= BF16_TOL
    # scalar-timestep form: one modulation row per group of rows
    per = (rows + r - 1) // r
    gi = (torch.arange(rows) // per).long()
    ref3 = (fp32_layer_norm(x.float(), None, None, 1e-6) * (1 + scale[gi]) + shift[gi]).bfloat16()
    out3 = ops.ln_modulate(x.cuda(), 1e-6, shift=tc[:, :dim], scale=tc[:, dim:2 * dim], rows_per_group=per)
    assert rel_err(out3, ref3) <= BF16_TOL


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("rows,dim", [(64, 1024), (1003, 3072), (4099, 3072), (777, 1160), (130, 5120), (70, 2048)])
def test_ln_kernel_variants_vs_oracle(ops, variant, rows, dim):
    """Each LayerNorm kernel (warp per row / block per row / batched block per row with the per-warp (sum, M2)
    combine) against the fp32 oracle: ragged row counts, a partial last warp (dim 1160), strided input, row groups
    that change inside a batch of rows."""
    from oracle.wan_oracle import fp32_layer_norm

    x = bf(rows, dim + 64, scale=2.0)
    x[:, 32:] += 3.0  # non-zero mean: the variance must not be computed as E[x^2] - mean^2
    xs = x.cuda()[:, 32:32 + dim]
    xf = x[:, 32:32 + dim].float()
    tab = torch.randn(3, 6 * dim, generator=torch.Generator().manual_seed(1)) * 0.5
    ridx = torch.randint(0, 3, (rows,), generator=torch.Generator().manual_seed(2)).int()
    shift, scale = tab[:, :dim], tab[:, dim:2 * dim]
    g, b = torch.randn(dim), torch.randn(dim)
    tc = tab.cuda()
    ops.rows_set_variant(variant, 2)
    try:
        out = ops.ln_modulate(xs, 1e-6, shift=tc[:, :dim], scale=tc[:, dim:2 * dim], row_index=ridx.cuda())
        ref = (fp32_layer_norm(xf, None, None, 1e-6) * (1 + scale[ridx.long()]) + shift[ridx.long()]).bfloat16()
        assert rel_err(out, ref) <= BF16_TOL
        out = ops.ln_modulate(xs, 1e-6, gamma=g.cuda(), beta=b.cuda())
        assert rel_err(out, fp32_layer_norm(xf, g, b, 1e-6).bfloat16()) <= BF16_TOL
        out = ops.ln_modulate(xs, 1e-6)
        assert rel_err(out, fp32_layer_norm(xf, None, None, 1e-6).bfloat16()) <= BF16_TOL
        per = 333
        gi = (torch.arange(rows) // per).long() % 3
        if rows <= 3 * per:
            out = ops.ln_modulate(xs, 1e-6, gamma=g.cuda(), beta=b.cuda(), shift=tc[:, :dim], scale=tc[:, dim:2 * dim],
                                  rows_per_group=per)
            ref = (fp32_layer_norm(xf, g, b, 1e-6) * (1 + scale[gi]) + shift[gi]).bfloat16()
            assert rel_err(out, ref) <= BF16_TOL
    finally:
        ops.rows_set_variant(3, 2)


def test_ln_bf16_module_flow_vs_torch(ops):
    """CogVideoXLayerNormZero body: bf16 LayerNorm, then bf16 (1+scale) multiply and shift add."""
    rows, dim = 300, 3072
    x = bf(rows, dim, scale=1.5)
    g, b = bf(dim, seed=3), bf(dim, seed=4)
    sc, sh = bf(2, dim, scale=0.3, seed=5), bf(2, dim, scale=0.3, seed=6)
    gi = (torch.arange(rows) // 150).long()
    ref = F.layer_norm(x, (dim,), g, b, 1e-5) * (1 + sc)[gi] + sh[gi]
    out = ops.ln_modulate(x.cuda(), 1e-5, gamma=g.float().cuda(), beta=b.float().cuda(), shift=sh.float().cuda(),
                          scale=sc.float().cuda(), rows_per_group=150, bf16_steps=True)
    assert rel_err(out, ref) <= BF16_TOL


def test_gate_residual(ops):
    x, y = bf(515, 768), bf(515, 768, seed=1)
    gate = torch.randn(2, 768)
    ridx = torch.randint(0, 2, (515,)).int()
    ref = (x.float() + y * gate[ridx.long()]).bfloat16()
    out = ops.gate_residual(x.cuda(), y.cuda(), gate.cuda(), row_index=ridx.cuda())
    assert rel_err(out, ref) <= BF16_TOL
    assert rel_err(ops.gate_residual(x.cuda(), y.cuda()), x + y) <= BF16_TOL


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("b,n,h,d", [(1, 384, 8, 32), (2, 300, 24, 128), (1, 77, 4, 128), (1, 45, 40, 128)])
def test_qk_rmsnorm_rope_vs_oracle(ops, variant, b, n, h, d, request):
    from oracle.wan_oracle import apply_wan_rope, rms_norm

    # q/k kernel variants: warp per row, block per token, packed warp per row (the default at dim 3072 / 5120)
    ops.rows_set_variant(3, variant)
    request.addfinalizer(lambda: ops.rows_set_variant(3, 2))
    dm = h * d
    qkv = bf(b, n, 3 * dm)
    wq, wk = (1 + 0.1 * torch.randn(dm)).bfloat16(), (1 + 0.1 * torch.randn(dm)).bfloat16()
    ang = torch.rand(n, d // 2) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).contiguous()
    sin = ang.sin().repeat_interleave(2, dim=1).contiguous()

    def ref(x, w):
        y = rms_norm(x, w, 1e-6).unflatten(2, (h, -1)).transpose(1, 2)
        y = apply_wan_rope(y, cos[None, None], sin[None, None])
        return y.transpose(1, 2).flatten(2)

    rq, rk = ref(qkv[..., :dm], wq), ref(qkv[..., dm:2 * dm], wk)
    g = qkv.cuda()
    ops.qk_norm_rope(g[..., :dm], wq.cuda(), g[..., dm:2 * dm], wk.cuda(), h, rope_mode=ops.ROPE_WAN, cos=cos.cuda(),
                     sin=sin.cuda(), seq_len=n)
    assert rel_err(g[..., :dm], rq) <= BF16_TOL
    assert rel_err(g[..., dm:2 * dm], rk) <= BF16_TOL
    assert torch.equal(g[..., 2 * dm:].cpu(), qkv[..., 2 * dm:])  # V untouched
    # cross-attention form: different row counts, no rope
    q2, k2 = bf(n, dm, seed=7), bf(16, dm, seed=8)
    gq, gk = q2.cuda(), k2.cuda()
    ops.qk_norm_rope(gq, wq.cuda(), gk, wk.cuda(), h)
    assert rel_err(gq, rms_norm(q2, wq, 1e-6)) <= BF16_TOL
    assert rel_err(gk, rms_norm(k2, wk, 1e-6)) <= BF16_TOL


def test_qk_layernorm_rope_cog_vs_oracle(ops):
    from oracle.cog_oracle import apply_cog_rope

    b, text, nv, h, d = 2, 10, 150, 6, 64
    s, dm = text + nv, h * d
    qkv = bf(b, s, 3 * dm)
    w, bb = (1 + 0.1 * torch.randn(d)).bfloat16(), (0.1 * torch.randn(d)).bfloat16()
    ang = torch.rand(nv, d // 2) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).contiguous()
    sin = ang.sin().repeat_interleave(2, dim=1).contiguous()

    def ref(x):
        y = F.layer_norm(x.view(b, s, h, d).transpose(1, 2), (d,), w, bb, 1e-6).clone()
        y[:, :, text:] = apply_cog_rope(y[:, :, text:], cos, sin)
        return y.transpose(1, 2).reshape(b, s, dm)

    rq, rk = ref(qkv[..., :dm]), ref(qkv[..., dm:2 * dm])
    g = qkv.cuda()
    ops.qk_norm_rope(g[..., :dm], w.cuda(), g[..., dm:2 * dm], w.cuda(), h, b0=bb.cuda(), b1=bb.cuda(),
                     norm_mode=ops.QK_LAYERNORM_PER_HEAD, rope_mode=ops.ROPE_COGVIDEOX, cos=cos.cuda(), sin=sin.cuda(),
                     seq_len=s, rope_skip=text)
    assert rel_err(g[..., :dm], rq) <= BF16_TOL
    assert rel_err(g[..., dm:2 * dm], rk) <= BF16_TOL


@pytest.mark.parametrize("m,n,k", [(1, 8, 8), (128, 256, 64), (300, 192, 384), (1000, 3072, 1024), (2, 18432, 512),
                                   (4097, 264, 136)])
def test_gemm_vs_oracle(ops, m, n, k):
    a = bf(m, k)
    w = bf(n, k, scale=1 / math.sqrt(k), seed=1)
    bias = bf(n, seed=2)
    ref = F.linear(a.float(), w.float(), bias.float())
    assert rel_err(ops.linear(a.cuda(), w.cuda(), bias.cuda()), ref) <= BF16_TOL
    assert rel_err(ops.linear(a.cuda(), w.cuda(), None), F.linear(a.float(), w.float())) <= BF16_TOL
    assert rel_err(ops.linear(a.cuda(), w.cuda(), bias.cuda(), out_dtype=torch.float32), ref) <= 1e-4


def test_gemm_epilogues_vs_oracle(ops):
    m, n, k = 777, 1024, 512
    a, w, bias = bf(m, k), bf(n, k, scale=1 / math.sqrt(k), seed=1), bf(n, seed=2)
    lin = F.linear(a.float(), w.float(), bias.float()).bfloat16()
    ag, wg, bg = a.cuda(), w.cuda(), bias.cuda()
    assert rel_err(ops.linear(ag, wg, bg, epilogue=ops.EPI_GELU_TANH), F.gelu(lin, approximate="tanh")) <= BF16_TOL
    assert rel_err(ops.linear(ag, wg, bg, epilogue=ops.EPI_SILU), F.silu(lin)) <= BF16_TOL
    x = bf(m, n, seed=3)
    gate = torch.randn(3, n)
    ridx = torch.randint(0, 3, (m,)).int()
    ref = (x.float() + lin * gate[ridx.long()]).bfloat16()  # transformer_wan.py:336
    out = ops.linear(ag, wg, bg, epilogue=ops.EPI_GATE_RESIDUAL, residual=x.cuda(), gate=gate.cuda(),
                     row_index=ridx.cuda())
    assert rel_err(out, ref) <= BF16_TOL
    xc = x.cuda().clone()  # in place, as the block uses it
    ops.linear(ag, wg, bg, epilogue=ops.EPI_GATE_RESIDUAL, residual=xc, gate=gate.cuda(), row_index=ridx.cuda(), out=xc)
    assert rel_err(xc, ref) <= BF16_TOL
    out = ops.linear(ag, wg, bg, epilogue=ops.EPI_GATE_RESIDUAL, residual=x.cuda())  # transformer_wan.py:341
    assert rel_err(out, x + lin) <= BF16_TOL
    gb = gate.bfloat16()
    ref_cog = x + gb[ridx.long()] * lin  # cogvideox_transformer_3d.py:146 (all bf16)
    out = ops.linear(ag, wg, bg, epilogue=ops.EPI_GATE_RESIDUAL, residual=x.cuda(), gate=gb.float().cuda(),
                     row_index=ridx.cuda(), round_product=True)
    assert rel_err(out, ref_cog) <= BF16_TOL


@pytest.mark.parametrize("splits", [2, 3])
def test_gemm_split_k_tail_matches_fused_epilogues(ops, splits):
    """The split-K work units of the pair kernel (raw fp32 slices + fix-up kernel with the fused epilogue) forced on
    the last round: every epilogue, ragged M and N, in-place gated residual, against the unsplit kernel and torch."""
    m, n, k = 700, 520, 1536
    a, w, b = bf(m, k), bf(n, k, scale=1.0 / math.sqrt(k), seed=1), bf(n, seed=2)
    x = bf(m, n, seed=3)
    gate = torch.randn(3, n, generator=torch.Generator().manual_seed(4))
    ridx = torch.randint(0, 3, (m,), generator=torch.Generator().manual_seed(5)).int()
    ac, wc, bc = a.cuda(), w.cuda(), b.cuda()
    y = F.linear(a.float(), w.float(), b.float())
    refs = {
        ops.EPI_NONE: y,
        ops.EPI_GELU_TANH: F.gelu(y.bfloat16().float(), approximate="tanh"),
        ops.EPI_SILU: F.silu(y.bfloat16().float()),
        ops.EPI_GATE_RESIDUAL: x.float() + y.bfloat16().float() * gate[ridx.long()],
    }
    try:
        for epi, ref in refs.items():
            kw = {}
            if epi == ops.EPI_GATE_RESIDUAL:
                kw = dict(residual=x.cuda(), gate=gate.cuda(), row_index=ridx.cuda())
            ops.gemm_set_split(0)
            whole = ops.linear(ac, wc, bc, epilogue=epi, **kw)
            ops.gemm_set_split(splits)
            if epi == ops.EPI_GATE_RESIDUAL:  # in place, as the blocks use it
                xs = x.cuda()
                out = ops.linear(ac, wc, bc, epilogue=epi, residual=xs, gate=gate.cuda(), row_index=ridx.cuda(), out=xs)
            else:
                out = ops.linear(ac, wc, bc, epilogue=epi, **kw)
            assert rel_err(out, ref) <= BF16_TOL, epi
            assert rel_err(out, whole.float().cpu()) <= 2.0 ** -7, epi  # same sums, different association
    finally:
        ops.gemm_set_split(-1)


def _sdpa_ref(q, k, v, heads):
    from oracle.wan_oracle import sdpa

    b, nq, inner = q.shape
    d = inner // heads
    o = sdpa(q.view(b, nq, heads, d).transpose(1, 2).float(), k.view(b, -1, heads, d).transpose(1, 2).float(),
             v.view(b, -1, heads, d).transpose(1, 2).float())
    return o.transpose(1, 2).reshape(b, nq, inner)


@pytest.mark.parametrize("hd", [128, 64])
@pytest.mark.parametrize("b,h,nq,nk", [(1, 1, 1, 1), (1, 1, 256, 128), (1, 2, 130, 257), (2, 3, 1000, 1000),
                                       (1, 2, 300, 16), (1, 2, 2048, 2048)])
def test_attention_vs_oracle(ops, hd, b, h, nq, nk):
    dm = h * hd
    q, k, v = bf(b, nq, dm), bf(b, nk, dm, seed=1), bf(b, nk, dm, seed=2)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
    assert rel_err(out, _sdpa_ref(q, k, v, h)) <= BF16_TOL


@pytest.mark.parametrize("hd", [128, 64])
@pytest.mark.parametrize("splits", [2, 3, 7])
def test_attention_kv_split_matches_oracle_and_single_pass(ops, hd, splits):
    """The KV-split decomposition (partial CTAs + merge kernel, used for the partly filled last wave) forced on every
    tile: ragged nq/nk, more splits than some shapes have KV tiles, peaky scores (different running maxima per split)."""
    try:
        for (b, h, nq, nk, scale) in [(1, 2, 300, 1000, 1.0), (2, 3, 513, 897, 2.5), (1, 1, 64, 129, 1.0)]:
            dm = h * hd
            q, k, v = bf(b, nq, dm, scale=scale), bf(b, nk, dm, scale=scale, seed=1), bf(b, nk, dm, seed=2)
            ops.attention_set_split(0)
            single = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
            ops.attention_set_split(splits)
            out = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
            assert rel_err(out, _sdpa_ref(q, k, v, h)) <= BF16_TOL
            assert rel_err(out, single.float().cpu()) <= 2.0 ** -7  # same math up to one bf16 rounding of the merge
    finally:
        ops.attention_set_split(-1)


@pytest.mark.parametrize("hd,variant", [(128, v) for v in range(1, 6)] + [(64, v) for v in range(1, 14)])
def test_attention_scheduling_variants_agree(ops, hd, variant):
    """Every scheduling variant behind fino_attention_set_variant (exponentials on MUFU vs FMA pipes, split P
    publication, separate P columns + early S issue and softmax turn-taking at head_dim 64) computes the same function:
    ragged shapes, peaky scores (lazy rescale), KV split on top."""
    try:
        ops.attention_set_variant(variant)
        for (b, h, nq, nk, scale, split) in [(1, 2, 300, 1000, 1.0, -1), (2, 3, 513, 897, 2.5, -1), (1, 1, 64, 129, 1.0, -1),
                                             (1, 2, 700, 2100, 2.5, 3)]:
            dm = h * hd
            q, k, v = bf(b, nq, dm, scale=scale), bf(b, nk, dm, scale=scale, seed=1), bf(b, nk, dm, seed=2)
            ops.attention_set_split(split)
            out = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
            assert rel_err(out, _sdpa_ref(q, k, v, h)) <= BF16_TOL, (variant, b, h, nq, nk)
    finally:
        ops.attention_set_variant(0)
        ops.attention_set_split(-1)


def test_attention_auto_split_on_a_partial_wave(ops):
    """Automatic mode on a shape whose tile count leaves a partly filled last wave (the 8-way Ulysses situation)."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    h, hd, nk = 1, 128, 2048
    nq = (sms + 5) * 256  # one full wave + 5 tiles
    n_full, s = ops.attention_plan(nq, nk, h, 1, sms)
    assert n_full == sms and s >= 2
    q, k, v = bf(1, nq, hd), bf(1, nk, hd, seed=1), bf(1, nk, hd, seed=2)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
    ops.attention_set_split(0)
    try:
        single = ops.attention(q.cuda(), k.cuda(), v.cuda(), h)
    finally:
        ops.attention_set_split(-1)
    assert torch.equal(out[:, : sms * 256], single[:, : sms * 256])  # whole tiles take the single-pass path
    assert rel_err(out[:, sms * 256:], single[:, sms * 256:].float().cpu()) <= 2.0 ** -7
    assert rel_err(out[:, -1500:], _sdpa_ref(q[:, -1500:].contiguous(), k, v, h)) <= BF16_TOL


@pytest.mark.parametrize("hd", [128, 64])
def test_attention_strided_fused_qkv_and_peaky_scores(ops, hd):
    b, h, n = 2, 3, 700
    dm = h * hd
    qkv = bf(b, n, 3 * dm, scale=2.5)  # |scores| up to ~70: exercises the lazy-rescale path of the online softmax
    g = qkv.cuda()
    out = ops.attention(g[..., :dm], g[..., dm:2 * dm], g[..., 2 * dm:], h)
    ref = _sdpa_ref(qkv[..., :dm].contiguous(), qkv[..., dm:2 * dm].contiguous(), qkv[..., 2 * dm:].contiguous(), h)
    assert rel_err(out, ref) <= BF16_TOL


def test_patchify_unpatchify_vs_oracle(ops):
    b, c, f, h, w, dm = 1, 32, 6, 16, 16, 256
    x = bf(b, c, f, h, w)
    wt, bias = bf(dm, c, 1, 2, 2, scale=0.1, seed=1), bf(dm, seed=2)
    ref = F.conv3d(x.float(), wt.float(), bias.float(), stride=(1, 2, 2)).flatten(2).transpose(1, 2).reshape(-1, dm)
    xg = x.cuda()
    rows = ops.patchify(xg, (b, c, f, h, w), xg.stride(), (1, 2, 2))
    y = ops.linear(rows, wt.cuda().view(dm, -1), bias.cuda())
    assert rel_err(y, ref) <= BF16_TOL
    # Wan unpatchify (transformer_wan.py:539-543)
    co = 16
    r = bf(b * f * (h // 2) * (w // 2), 4 * co, seed=3)
    ref = r.reshape(b, f, h // 2, w // 2, 1, 2, 2, -1).permute(0, 7, 1, 4, 2, 5, 3, 6).flatten(6, 7).flatten(4, 5).flatten(2, 3)
    out = torch.empty(b, co, f, h, w, dtype=torch.bfloat16, device="cuda")
    ops.unpatchify(r.cuda(), out, (b, co, f, h, w), out.stride(), (1, 2, 2), True)
    assert torch.equal(out.cpu(), ref)
    # CogVideoX unpatchify (cogvideox_transformer_3d.py:549-550), output layout [B, F, C, H, W]
    r = bf(2 * 3 * 4 * 6, 16 * 4, seed=4)
    ref = r.reshape(2, 3, 4, 6, -1, 2, 2).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    out = torch.empty(2, 3, 16, 8, 12, dtype=torch.bfloat16, device="cuda")
    st = out.stride()
    ops.unpatchify(r.cuda(), out, (2, 16, 3, 8, 12), (st[0], st[2], st[1], st[3], st[4]), (1, 2, 2), False)
    assert torch.equal(out.cpu(), ref)


def test_timestep_embedding_and_small_linear_vs_oracle(ops):
    from oracle.wan_oracle import sinusoidal_embedding

    t = torch.tensor([0.0, 500.0, 999.0, 37.5])
    ref = sinusoidal_embedding(t, 256)
    out = ops.timestep_embedding(t.cuda(), 256)
    assert float((out.cpu() - ref).abs().max()) <= 2e-4  # sin/cos of arguments up to 1e3 in fp32
    x = torch.randn(2, 256)
    w, b = torch.randn(3072, 256) * 0.05, torch.randn(3072) * 0.1
    ref = F.silu(F.linear(x, w, b))
    assert rel_err(ops.linear_small_m(x.cuda(), w.cuda(), b.cuda(), act_out=1), ref) <= 1e-5
    wb, bb = w.bfloat16(), b.bfloat16()
    ref = F.linear(F.silu(x).bfloat16(), wb, bb)
    out = ops.linear_small_m(x.cuda(), wb.cuda(), bb.cuda(), act_in=1, round_in=True, round_out=True)
    assert rel_err(out, ref) <= BF16_TOL


def test_swap01_and_mod_table(ops):
    x = bf(6, 4, 64)
    out = ops.swap01(x.cuda().view(-1), 6, 4, 64)
    assert torch.equal(out.cpu(), x.transpose(0, 1).contiguous())
    tab, proj = torch.randn(5, 6 * 64), torch.randn(3, 6 * 64)
    o = ops.build_mod_table(tab.cuda(), proj.cuda(), 5, 6 * 64)
    assert torch.equal(o.cpu(), tab[:, None] + proj[None])


def test_invalid_arguments_raise(ops):
    from frameino_b200._lib import FinoError

    with pytest.raises(FinoError):
        ops.attention(bf(1, 8, 136).cuda(), bf(1, 8, 136).cuda(), bf(1, 8, 136).cuda(), 1)  # head_dim 136 unsupported
    with pytest.raises(FinoError):
        ops.linear(bf(4, 12).cuda(), bf(8, 12).cuda())  # K not a multiple of 8


@pytest.mark.parametrize("rows,dim", [(1003, 3072), (256, 1024), (2049, 4096), (777, 2048)])
def test_rows_tma_kernels_match_warp_kernels(ops, rows, dim):
    """The TMA-staged persistent row kernels (rows_tma.cu) against the register-resident warp-per-row kernels on the
    same inputs: ragged row counts (tail stage), strided views, both modulation-row selectors, q-only and q+k forms."""
    heads = dim // 128
    x = bf(rows, dim + 64, scale=2.0)
    tab = torch.randn(3, 6 * dim, generator=torch.Generator().manual_seed(1)).cuda() * 0.5
    ridx = torch.randint(0, 3, (rows,), generator=torch.Generator().manual_seed(2)).int().cuda()
    g, b = torch.randn(dim).cuda(), torch.randn(dim).cuda()
    xc = x.cuda()[:, 32:32 + dim]  # strided view: per-row bulk copies
    assert not xc.is_contiguous()
    qkv = bf(1, rows, 3 * dim, seed=4).cuda()
    wq, wk = (1 + 0.1 * torch.randn(dim)).bfloat16().cuda(), (1 + 0.1 * torch.randn(dim)).bfloat16().cuda()
    ang = torch.rand(rows, 64) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=1).contiguous().cuda()
    sin = ang.sin().repeat_interleave(2, dim=1).contiguous().cuda()
    res = {}
    for tma in (False, True):
        ops.rows_set_tma(tma)
        try:
            r = []
            r.append(ops.ln_modulate(xc, 1e-6, shift=tab[:, :dim], scale=tab[:, dim:2 * dim], row_index=ridx))
            r.append(ops.ln_modulate(xc.contiguous(), 1e-6, shift=tab[:, :dim], scale=tab[:, dim:2 * dim],
                                     rows_per_group=(rows + 2) // 3))
            r.append(ops.ln_modulate(xc, 1e-6, gamma=g, beta=b))
            r.append(ops.ln_modulate(xc.contiguous(), 1e-5, gamma=g, beta=b, shift=tab[:, :dim],
                                     scale=tab[:, dim:2 * dim], row_index=ridx, bf16_steps=True))
            t = qkv.clone()
            ops.qk_norm_rope(t[..., :dim], wq, t[..., dim:2 * dim], wk, heads, rope_mode=ops.ROPE_WAN, cos=cos, sin=sin,
                             seq_len=rows)
            r.append(t)
            t2 = qkv.clone()
            ops.qk_norm_rope(t2[..., :dim], wq, t2[..., dim:2 * dim], wk, heads)  # no rope
            r.append(t2)
            q3, k3 = qkv[0, :, :dim].contiguous(), qkv[0, :16, dim:2 * dim].contiguous()
            ops.qk_norm_rope(q3, wq, k3, wk, heads)  # cross-attention form
            r += [q3, k3]
            res[tma] = r
        finally:
            ops.rows_set_tma(False)
    for i, (a, c) in enumerate(zip(res[False], res[True])):
        if i < 4:  # LayerNorm: the block-wide fp32 sums associate differently -> last-bit differences only
            assert rel_err(c, a) <= 4e-3, i
        else:      # RMSNorm + RoPE: one rsqrt per row, everything after it is per element -> bf16-ulp agreement
            assert rel_err(c, a) <= 8e-3, i  # <= 2 bf16 ulps at the largest magnitude
