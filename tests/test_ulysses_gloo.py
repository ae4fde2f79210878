"""World-size-2 (and 4) gloo run of the Ulysses exchange on CPU: the pack / all-to-all / unpack layout math of
frameino_b200/ulysses.py against un-sharded attention, with torch stand-ins for the two CUDA primitives."""
import os
import socket

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


def _torch_swap01(x, a, b, inner, out=None):
    return x.view(a, b, inner).transpose(0, 1).contiguous()


def _torch_attention(q, k, v, heads, scale=None, out=None):
    b, nq, inner = q.shape
    d = inner // heads
    qh = q.reshape(b, nq, heads, d).transpose(1, 2).float()
    kh = k.reshape(b, -1, heads, d).transpose(1, 2).float()
    vh = v.reshape(b, -1, heads, d).transpose(1, 2).float()
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, scale=scale)
    return o.transpose(1, 2).reshape(b, nq, inner).to(q.dtype)


def _worker(rank, world, port, n_total, heads, hd, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frameino_b200.ulysses import SequenceParallel

        sp = SequenceParallel()
        sp._swap01 = _torch_swap01
        sp._attention = _torch_attention
        sp.plan(n_total)
        torch.manual_seed(0)
        d_model = heads * hd
        qkv_full = torch.randn(1, n_total, 3 * d_model)
        ref = _torch_attention(qkv_full[..., :d_model], qkv_full[..., d_model:2 * d_model], qkv_full[..., 2 * d_model:],
                               heads, scale=hd ** -0.5)
        local = sp.shard_rows(qkv_full)
        assert local.shape == (1, sp.n_loc, 3 * d_model)
        out = sp.attention(local, heads, hd ** -0.5)
        sl = sp.local_slice()
        n_real = sl.stop - sl.start
        err = float((out[:, :n_real] - ref[:, sl]).abs().max()) if n_real > 0 else 0.0
        gathered = sp.gather_rows(out)
        err_g = float((gathered - ref).abs().max())
        ret[rank] = (err, err_g, tuple(gathered.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,heads", [(2, 64, 4), (2, 37, 2), (4, 50, 8)])
def test_ulysses_exchange_matches_full_attention(world, n_total, heads):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_total, heads, 16, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, err_g, shape = ret[rank]
        assert err < 1e-5, (rank, err)
        assert err_g < 1e-5, (rank, err_g)
        assert shape == (1, n_total, heads * 16)
