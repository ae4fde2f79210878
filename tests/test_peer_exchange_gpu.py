"""-m gpu: the fused Ulysses exchange kernels (frameino_b200/csrc/peer_kernels.cu + the owner-scatter attention
epilogue), driven through the C ABI on ONE GPU. P ranks are emulated in one process: every rank gets its own
``fino_peer_alloc`` buffer (all local here, peer-mapped in production) and the pointer tables are built by the very
``exchange_layout`` arithmetic production uses. Compared bit for bit with the un-sharded kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import ops as _ops

    return _ops


def _inputs(n_total, heads, hd, seed=0):
    g = torch.Generator().manual_seed(seed)
    d = heads * hd
    qkv = (torch.randn(1, n_total, 3 * d, generator=g) * 1.5).bfloat16().cuda()
    wq = (1 + 0.1 * torch.randn(d, generator=g)).bfloat16().cuda()
    wk = (1 + 0.1 * torch.randn(d, generator=g)).bfloat16().cuda()
    ang = torch.rand(n_total, hd // 2, generator=g) * 6.28
    cos = ang.cos().repeat_interleave(2, 1).contiguous().cuda()
    sin = ang.sin().repeat_interleave(2, 1).contiguous().cuda()
    return qkv, wq, wk, cos, sin


@pytest.mark.parametrize("world,n_total,heads,hd,tma", [(1, 300, 2, 128, False), (2, 1024, 4, 128, False),
                                                       (4, 777, 4, 128, False), (8, 1000, 8, 64, False),
                                                       (2, 515, 6, 64, False), (2, 1201, 8, 128, True),
                                                       (8, 1003, 24, 128, False), (4, 520, 24, 128, False),
                                                       (2, 300, 24, 128, False), (8, 261, 40, 128, False),
                                                       (4, 2055, 16, 64, True)])
def test_fused_exchange_matches_unsharded(ops, world, n_total, heads, hd, tma):
    try:
        ops.attention_set_split(0)  # bit-exact comparison: keep the KV-split merge out of both sides
        _run_fused_exchange(ops, world, n_total, heads, hd, tma)
    finally:
        ops.rows_set_tma(False)
        ops.attention_set_split(-1)


def _same(a, b, exact):
    if exact:
        return torch.equal(a, b)
    # TMA-staged scatter kernel: its block-wide sum of squares associates differently -> <= 2 bf16 ulps
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9)) <= 1e-2


def _run_fused_exchange(ops, world, n_total, heads, hd, tma):
    from frameino_b200.ulysses import SequenceParallel, exchange_layout

    d = heads * hd
    qkv, wq, wk, cos, sin = _inputs(n_total, heads, hd)
    # un-sharded reference: in-place norm+rope then attention, same kernels' single-GPU forms
    ref_qkv = qkv.clone()
    ops.rows_set_tma(False)
    ops.qk_norm_rope(ref_qkv[..., :d], wq, ref_qkv[..., d:2 * d], wk, heads, norm_mode=ops.QK_RMS_ACROSS_HEADS,
                     eps=1e-6, rope_mode=ops.ROPE_WAN, cos=cos, sin=sin, seq_len=n_total)
    ref = ops.attention(ref_qkv[..., :d], ref_qkv[..., d:2 * d], ref_qkv[..., 2 * d:], heads)
    ops.rows_set_tma(tma)  # True: the TMA-staged scatter kernel (eligible configs only); the reference stays un-staged
    exact = not tma

    n_loc, n_pad = SequenceParallel.partition(n_total, world)
    lays = [exchange_layout(world, r, n_loc, d) for r in range(world)]
    bases = [ops.peer_alloc(lays[0]["total_bytes"]) for _ in range(world)]
    try:
        inner = lays[0]["inner"]
        qkv_ptrs = ops.pointer_table([b + lays[0]["qkv_off"] for b in bases])
        flag_ptrs = ops.pointer_table([b + lays[0]["flags_off"] for b in bases])
        # phase 1: every rank normalises + rotates its token slice and scatters head groups to the owners
        for r in range(world):
            lo, hi = r * n_loc, min((r + 1) * n_loc, n_total)
            local = torch.zeros(1, n_loc, 3 * d, dtype=torch.bfloat16, device="cuda")
            lcos = torch.zeros(n_loc, hd, device="cuda")
            lsin = torch.zeros(n_loc, hd, device="cuda")
            if hi > lo:
                local[:, : hi - lo] = qkv[:, lo:hi]
                lcos[: hi - lo] = cos[lo:hi]
                lsin[: hi - lo] = sin[lo:hi]
            ops.qkv_norm_rope_scatter(local, wq, wk, heads, 1e-6, lcos, lsin, qkv_ptrs, world, r, n_loc,
                                      lays[r]["qkv_row_stride"])
        # every rank now holds all tokens of its head group == the matching columns of the reference
        for r in range(world):
            got = ops.tensor_from_ptr(bases[r] + lays[r]["qkv_off"], (1, n_pad, 3 * inner))[:, :n_total]
            for part in range(3):
                want = ref_qkv[..., part * d + r * inner: part * d + (r + 1) * inner]
                assert _same(got[..., part * inner:(part + 1) * inner], want, exact or part == 2), (r, part)
        # the barrier kernel with one participant per emulated rank, each on its own stream (all must be resident)
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()
        for epoch in (1, 2):
            for r, st in enumerate(streams):
                with torch.cuda.stream(st):
                    ops.peer_barrier(flag_ptrs, r, world, epoch)
        torch.cuda.synchronize()
        # phase 2: attention on each rank's heads, output rows scattered to the rows' owners
        for r in range(world):
            full = ops.tensor_from_ptr(bases[r] + lays[r]["qkv_off"], (1, n_pad, 3 * inner))[:, :n_total]
            o_ptrs = ops.pointer_table([b + lays[r]["o_off"] + lays[r]["o_col_offset"] for b in bases])
            ops.attention_scatter(full[..., :inner], full[..., inner:2 * inner], full[..., 2 * inner:], heads // world,
                                  o_ptrs, world, n_loc, lays[r]["o_row_stride"])
        torch.cuda.synchronize()
        for r in range(world):
            lo, hi = r * n_loc, min((r + 1) * n_loc, n_total)
            o_loc = ops.tensor_from_ptr(bases[r] + lays[r]["o_off"], (1, n_loc, d))
            if hi > lo:
                assert _same(o_loc[:, : hi - lo], ref[:, lo:hi], exact), r
            if hi - lo < n_loc:  # pad rows are never written: still the zero fill of peer_alloc
                assert not o_loc[:, max(hi - lo, 0):].any()
    finally:
        torch.cuda.synchronize()
        for b in bases:
            ops.peer_free(b)


@pytest.mark.parametrize("world,n_total,heads,text_len,batch", [(2, 530, 4, 10, 2), (4, 777, 8, 190, 1), (8, 1001, 48, 100, 2),
                                                                (1, 300, 4, 0, 1)])
def test_fused_exchange_cogvideox_matches_unsharded(ops, world, n_total, heads, text_len, batch):
    """The CogVideoX prologue (per-head LayerNorm(64) + RoPE on the video rows of the JOINT sequence) fused with the
    first exchange, batch slots of the exchange buffer, then the attention + second exchange; against the un-sharded
    kernels on the same data."""
    from frameino_b200.ulysses import SequenceParallel, exchange_layout

    hd = 64
    d = heads * hd
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(batch, n_total, 3 * d, generator=g) * 1.5).bfloat16().cuda()
    wq, bq, wk, bk = ((1 + 0.1 * torch.randn(hd, generator=g)).bfloat16().cuda() for _ in range(4))
    ang = torch.rand(n_total - text_len, hd // 2, generator=g) * 6.28
    cos = ang.cos().repeat_interleave(2, 1).contiguous().cuda()
    sin = ang.sin().repeat_interleave(2, 1).contiguous().cuda()
    ops.attention_set_split(0)
    try:
        ref_qkv = qkv.clone()
        ops.qk_norm_rope(ref_qkv[..., :d], wq, ref_qkv[..., d:2 * d], wk, heads, b0=bq, b1=bk,
                         norm_mode=ops.QK_LAYERNORM_PER_HEAD, eps=1e-6, rope_mode=ops.ROPE_COGVIDEOX, cos=cos, sin=sin,
                         seq_len=n_total, rope_skip=text_len)
        ref = ops.attention(ref_qkv[..., :d], ref_qkv[..., d:2 * d], ref_qkv[..., 2 * d:], heads)
        n_loc, n_pad = SequenceParallel.partition(n_total, world)
        assert text_len <= n_loc
        lays = [exchange_layout(world, r, n_loc, d, batch=batch) for r in range(world)]
        bases = [ops.peer_alloc(lays[0]["total_bytes"]) for _ in range(world)]
        try:
            inner = lays[0]["inner"]
            qb, ob = lays[0]["qkv_batch_bytes"], lays[0]["o_batch_bytes"]
            for bi in range(batch):
                qkv_ptrs = ops.pointer_table([b + lays[0]["qkv_off"] + bi * qb for b in bases])
                for r in range(world):
                    lo, hi = r * n_loc, min((r + 1) * n_loc, n_total)
                    skip = text_len if r == 0 else 0
                    local = torch.zeros(n_loc, 3 * d, dtype=torch.bfloat16, device="cuda")
                    lcos = torch.zeros(n_loc - skip, hd, device="cuda")
                    lsin = torch.zeros(n_loc - skip, hd, device="cuda")
                    if hi > lo:
                        local[: hi - lo] = qkv[bi, lo:hi]
                        v0 = lo + skip - text_len  # first video row of this rank
                        lcos[: hi - lo - skip] = cos[v0:v0 + hi - lo - skip]
                        lsin[: hi - lo - skip] = sin[v0:v0 + hi - lo - skip]
                    ops.qkv_ln_rope_scatter(local, wq, bq, wk, bk, heads, 1e-6, lcos, lsin, skip, qkv_ptrs, world, r, n_loc,
                                            lays[r]["qkv_row_stride"])
                for r in range(world):
                    got = ops.tensor_from_ptr(bases[r] + lays[r]["qkv_off"] + bi * qb, (1, n_pad, 3 * inner))[:, :n_total]
                    for part in range(3):
                        want = ref_qkv[bi:bi + 1, :, part * d + r * inner: part * d + (r + 1) * inner]
                        assert _same(got[..., part * inner:(part + 1) * inner], want, part == 2), (bi, r, part)
                for r in range(world):
                    full = ops.tensor_from_ptr(bases[r] + lays[r]["qkv_off"] + bi * qb, (1, n_pad, 3 * inner))[:, :n_total]
                    o_ptrs = ops.pointer_table([b + lays[r]["o_off"] + bi * ob + lays[r]["o_col_offset"] for b in bases])
                    ops.attention_scatter(full[..., :inner], full[..., inner:2 * inner], full[..., 2 * inner:],
                                          heads // world, o_ptrs, world, n_loc, lays[r]["o_row_stride"])
            torch.cuda.synchronize()
            for bi in range(batch):
                for r in range(world):
                    lo, hi = r * n_loc, min((r + 1) * n_loc, n_total)
                    o_loc = ops.tensor_from_ptr(bases[r] + lays[r]["o_off"] + bi * ob, (1, n_loc, d))
                    if hi > lo:
                        assert _same(o_loc[:, : hi - lo], ref[bi:bi + 1, lo:hi], False), (bi, r)
        finally:
            torch.cuda.synchronize()
            for b in bases:
                ops.peer_free(b)
    finally:
        ops.attention_set_split(-1)


def test_ipc_export_import_roundtrip_same_process(ops):
    """cudaIpcGetMemHandle works on a peer_alloc buffer (opening it needs a second process: covered by
    tools/sp_check.py on a multi-GPU box)."""
    p = ops.peer_alloc(4096)
    try:
        h = ops.peer_export(p)
        assert isinstance(h, bytes) and len(h) == 64 and any(h)
        t = ops.tensor_from_ptr(p, (2048,), torch.bfloat16)
        assert t.is_cuda and not t.any()
        t.fill_(1.0)
        assert float(ops.tensor_from_ptr(p, (2048,), torch.bfloat16).sum()) == 2048.0
    finally:
        ops.peer_free(p)


def test_attention_epilogue_ragged_and_strided(ops):
    """The shared-memory staged epilogue: ragged last tile, strided output view, d = 64 and 128."""
    from oracle.wan_oracle import sdpa

    for heads, hd, nq, nk in [(3, 128, 333, 200), (5, 64, 130, 515)]:
        d = heads * hd
        g = torch.Generator().manual_seed(5)
        q = torch.randn(2, nq, d, generator=g).bfloat16()
        k = torch.randn(2, nk, d, generator=g).bfloat16()
        v = torch.randn(2, nk, d, generator=g).bfloat16()
        out_buf = torch.zeros(2, nq, d + 64, dtype=torch.bfloat16, device="cuda")
        out = out_buf[..., 32:32 + d]
        ops.attention(q.cuda(), k.cuda(), v.cuda(), heads, out=out)
        qh, kh, vh = (t.float().view(2, -1, heads, hd).transpose(1, 2) for t in (q, k, v))
        ref = sdpa(qh, kh, vh).transpose(1, 2).reshape(2, nq, d)
        err = float((out.float().cpu() - ref).abs().max() / ref.abs().max())
        assert err <= 1e-2, (heads, hd, err)
        assert not out_buf[..., :32].any() and not out_buf[..., 32 + d:].any()
