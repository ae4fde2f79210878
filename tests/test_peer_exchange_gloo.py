"""World-size-2 / 4 gloo run (CPU) of the FUSED Ulysses path's host logic: ``exchange_layout`` arithmetic,
``PeerExchange`` handle swap + pointer tables, the scatter/barrier/attention/barrier sequence of
``SequenceParallel.fused_attention``. The device primitives are replaced by stand-ins that treat POSIX shared memory
as "peer memory" (pointer = mapped address, IPC handle = segment name), so the real cross-process addressing is
exercised without a GPU; the CUDA kernels themselves are covered by tests/test_peer_exchange_gpu.py."""
import ctypes
import os
import socket
from multiprocessing import shared_memory

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class ShmPrims:
    """Stand-ins for ops.peer_* / pointer_table / tensor_from_ptr on CPU."""

    def __init__(self):
        self.segments = {}

    @staticmethod
    def _addr(shm):
        return ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))

    def peer_alloc(self, nbytes):
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        shm.buf[:nbytes] = bytes(nbytes)
        a = self._addr(shm)
        self.segments[a] = (shm, True)
        return a

    def peer_export(self, ptr):
        return self.segments[ptr][0].name.encode().ljust(64, b"\0")

    def peer_import(self, handle):
        shm = shared_memory.SharedMemory(name=handle.rstrip(b"\0").decode())
        a = self._addr(shm)
        self.segments[a] = (shm, False)
        return a

    def peer_release(self, ptr):
        self.segments.pop(ptr)  # mapping stays alive until process exit (exported ctypes views pin it)

    def peer_free(self, ptr):
        shm, _ = self.segments.pop(ptr)
        try:
            shm.unlink()
        except FileNotFoundError:
            pass

    @staticmethod
    def pointer_table(ptrs):
        return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])

    @staticmethod
    def tensor_from_ptr(ptr, shape, dtype=torch.bfloat16):
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty(0, dtype=dtype).element_size()
        buf = (ctypes.c_char * nbytes).from_address(ptr)
        return torch.frombuffer(buf, dtype=dtype).view(*shape)

    @staticmethod
    def peer_barrier(flag_ptrs, rank, world, epoch):
        # mirrors the kernel: publish my epoch in every rank's flag array, then wait for everybody's
        for t in range(world):
            ctypes.c_uint32.from_address(flag_ptrs[t] + 4 * rank).value = epoch
        import time

        mine = flag_ptrs[rank]
        t0 = time.time()
        while any(ctypes.c_uint32.from_address(mine + 4 * t).value < epoch for t in range(world)):
            assert time.time() - t0 < 60, "barrier timeout"
            time.sleep(0.001)


def _rms(x, w, eps=1e-6):
    xf = x.float()
    return ((xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float() * w.float()).bfloat16()


def _torch_attention(q, k, v, heads, scale):
    b, nq, inner = q.shape
    d = inner // heads
    qh, kh, vh = (t.reshape(b, -1, heads, d).transpose(1, 2).float() for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    return o.transpose(1, 2).reshape(b, nq, inner).bfloat16()


def _worker(rank, world, port, n_total, heads, hd, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frameino_b200.ulysses import SequenceParallel

        prims = ShmPrims()
        sp = SequenceParallel(mode="peer")
        sp._prims = prims
        d_model = heads * hd

        def qkv_scatter(qkv, wq, wk, heads_, eps, cos, sin, dst_ptrs, world_, rank_, rows_per_rank, dst_row_stride):
            # what fino_qkv_norm_rope_scatter does (no RoPE in this host-logic test)
            n_loc = qkv.shape[1]
            inner = d_model // world_
            q = _rms(qkv[0, :, :d_model], wq, eps)
            k = _rms(qkv[0, :, d_model:2 * d_model], wk, eps)
            v = qkv[0, :, 2 * d_model:]
            for g in range(world_):
                dst = prims.tensor_from_ptr(dst_ptrs[g], (world_ * rows_per_rank, dst_row_stride))
                rows = slice(rank_ * rows_per_rank, rank_ * rows_per_rank + n_loc)
                for part, src in enumerate((q, k, v)):
                    dst[rows, part * inner:(part + 1) * inner] = src[:, g * inner:(g + 1) * inner]

        def attention_scatter(q, k, v, heads_, o_ptrs, num_owners, rows_per_owner, o_row_stride, scale):
            o = _torch_attention(q, k, v, heads_, scale)[0]
            inner = o.shape[1]
            for g in range(num_owners):
                lo, hi = g * rows_per_owner, min((g + 1) * rows_per_owner, o.shape[0])
                if hi <= lo:
                    continue
                # o_ptrs[g] already points at this rank's column block inside owner g's [n_loc, D] buffer
                flat = prims.tensor_from_ptr(o_ptrs[g], ((rows_per_owner - 1) * o_row_stride + inner,))
                flat.as_strided((hi - lo, inner), (o_row_stride, 1)).copy_(o[lo:hi])

        sp._qkv_scatter = qkv_scatter
        sp._attention_scatter = attention_scatter
        sp.plan(n_total)
        g = torch.Generator().manual_seed(0)
        qkv_full = torch.randn(1, n_total, 3 * d_model, generator=g).bfloat16()
        wq = (1 + 0.1 * torch.randn(d_model, generator=g)).bfloat16()
        wk = (1 + 0.1 * torch.randn(d_model, generator=g)).bfloat16()
        qn = _rms(qkv_full[..., :d_model], wq)
        kn = _rms(qkv_full[..., d_model:2 * d_model], wk)
        ref = _torch_attention(qn, kn, qkv_full[..., 2 * d_model:], heads, hd ** -0.5)
        local = sp.shard_rows(qkv_full)
        errs = []
        for _ in range(3):  # repeated layers reuse the buffers; the two barriers per call keep them race free
            out = sp.fused_attention(local, wq, wk, heads, 1e-6, None, None, hd ** -0.5)
            sl = sp.local_slice()
            n_real = max(sl.stop - sl.start, 0)
            errs.append(float((out[:, :n_real].float() - ref[:, sl].float()).abs().max()) if n_real else 0.0)
        lay = sp.exchange(sp.n_loc, d_model).layout
        ret[rank] = (max(errs), lay["o_col_offset"], sp.exchange(sp.n_loc, d_model).epoch)
        sp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,heads", [(2, 64, 4), (2, 37, 2), (4, 50, 8)])
def test_fused_exchange_host_logic(world, n_total, heads):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, heads, 16, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, col_off, epoch = ret[rank]
        assert err == 0.0, (rank, err)  # same bf16 arithmetic on both sides: the exchange itself is exact
        assert col_off == rank * (heads * 16 // world) * 2
        assert epoch == 6  # two barriers per fused attention


def test_exchange_layout_arithmetic():
    from frameino_b200.ulysses import exchange_layout

    for world, n_loc, d in [(1, 28160, 3072), (2, 14080, 3072), (8, 3520, 3072), (8, 4992, 3072), (4, 13, 512)]:
        lays = [exchange_layout(world, r, n_loc, d) for r in range(world)]
        inner = d // world
        for r, lay in enumerate(lays):
            assert lay["inner"] == inner and lay["n_pad"] == world * n_loc
            assert lay["qkv_off"] % 256 == 0 and lay["o_off"] % 256 == 0 and lay["total_bytes"] % 256 == 0
            assert lay["o_off"] >= lay["qkv_off"] + world * n_loc * 3 * inner * 2  # no overlap
            assert lay["total_bytes"] >= lay["o_off"] + n_loc * d * 2
            assert lay["o_col_offset"] == r * inner * 2
            # everything but the column offset is identical on every rank (peers index each other's buffers with it)
            assert {k: v for k, v in lay.items() if k != "o_col_offset"} == \
                   {k: v for k, v in lays[0].items() if k != "o_col_offset"}


# ---- CogVideoX form: per-head LayerNorm prologue, batched-CFG pair through one exchange (one slot per sample) --------
def _ln64(x, w, b, heads, eps=1e-6):
    n, d = x.shape
    y = torch.nn.functional.layer_norm(x.float().view(n, heads, 64), (64,), w.float(), b.float(), eps)
    return y.view(n, d).bfloat16()


def _cog_worker(rank, world, port, n_total, heads, batch, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frameino_b200.ulysses import SequenceParallel

        prims = ShmPrims()
        sp = SequenceParallel(mode="peer")
        sp._prims = prims
        d_model = heads * 64

        def ln_scatter(qkv, wq, bq, wk, bk, heads_, eps, cos, sin, rope_skip, dst_ptrs, world_, rank_, rows_per_rank,
                       dst_row_stride):
            # what fino_qkv_ln_rope_scatter does (no RoPE in this host-logic test); qkv: [n_loc, 3*D]
            n_loc = qkv.shape[0]
            inner = d_model // world_
            q = _ln64(qkv[:, :d_model], wq, bq, heads_, eps)
            k = _ln64(qkv[:, d_model:2 * d_model], wk, bk, heads_, eps)
            v = qkv[:, 2 * d_model:]
            for g in range(world_):
                dst = prims.tensor_from_ptr(dst_ptrs[g], (world_ * rows_per_rank, dst_row_stride))
                rows = slice(rank_ * rows_per_rank, rank_ * rows_per_rank + n_loc)
                for part, src in enumerate((q, k, v)):
                    dst[rows, part * inner:(part + 1) * inner] = src[:, g * inner:(g + 1) * inner]

        def attention_scatter(q, k, v, heads_, o_ptrs, num_owners, rows_per_owner, o_row_stride, scale):
            o = _torch_attention(q, k, v, heads_, scale)[0]
            inner = o.shape[1]
            for g in range(num_owners):
                lo, hi = g * rows_per_owner, min((g + 1) * rows_per_owner, o.shape[0])
                if hi <= lo:
                    continue
                flat = prims.tensor_from_ptr(o_ptrs[g], ((rows_per_owner - 1) * o_row_stride + inner,))
                flat.as_strided((hi - lo, inner), (o_row_stride, 1)).copy_(o[lo:hi])

        sp._qkv_ln_scatter = ln_scatter
        sp._attention_scatter = attention_scatter
        sp.plan(n_total)
        g = torch.Generator().manual_seed(1)
        qkv_full = torch.randn(batch, n_total, 3 * d_model, generator=g).bfloat16()

        class Norm:
            pass

        nq, nk = Norm(), Norm()
        nq.weight, nq.bias = (1 + 0.1 * torch.randn(64, generator=g)).bfloat16(), (0.1 * torch.randn(64, generator=g)).bfloat16()
        nk.weight, nk.bias = (1 + 0.1 * torch.randn(64, generator=g)).bfloat16(), (0.1 * torch.randn(64, generator=g)).bfloat16()
        refs = []
        for b in range(batch):
            qn = _ln64(qkv_full[b, :, :d_model], nq.weight, nq.bias, heads)
            kn = _ln64(qkv_full[b, :, d_model:2 * d_model], nk.weight, nk.bias, heads)
            refs.append(_torch_attention(qn[None], kn[None], qkv_full[b:b + 1, :, 2 * d_model:], heads, 0.125))
        ref = torch.cat(refs, dim=0)
        local = sp.shard_rows(qkv_full)
        errs = []
        for _ in range(2):
            out = sp.fused_attention_ln(local, nq, nk, heads, 1e-6, None, None, 0, 0.125)
            assert out.shape == (batch, sp.n_loc, d_model)
            sl = sp.local_slice()
            n_real = max(sl.stop - sl.start, 0)
            errs.append(float((out[:, :n_real].float() - ref[:, sl].float()).abs().max()) if n_real else 0.0)
        ex = sp.exchange(sp.n_loc, d_model, batch=batch)
        ret[rank] = (max(errs), ex.epoch, ex.layout["batch"], len(ex.qkv_ptrs_b))
        sp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,heads,batch", [(2, 41, 2, 2), (4, 30, 4, 1)])
def test_fused_exchange_cogvideox_host_logic(world, n_total, heads, batch):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cog_worker, args=(world, port, n_total, heads, batch, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, epoch, lay_batch, slots = ret[rank]
        assert err == 0.0, (rank, err)
        assert epoch == 4  # two barriers per call, whatever the batch
        assert lay_batch == batch and slots == batch


def test_exchange_layout_batch_slots():
    from frameino_b200.ulysses import exchange_layout

    one = exchange_layout(8, 3, 2391, 3072)
    two = exchange_layout(8, 3, 2391, 3072, batch=2)
    assert two["qkv_off"] == one["qkv_off"] and two["qkv_batch_bytes"] == one["qkv_batch_bytes"]
    assert two["qkv_batch_bytes"] % 256 == 0 and two["o_batch_bytes"] % 256 == 0
    assert two["o_off"] == 256 + 2 * two["qkv_batch_bytes"]
    assert two["total_bytes"] == two["o_off"] + 2 * two["o_batch_bytes"]
    assert two["qkv_batch_bytes"] >= 8 * 2391 * 3 * 384 * 2 and two["o_batch_bytes"] >= 2391 * 3072 * 2
