"""World-size-2 / 3 gloo run of the VAE's row-parallel plumbing on CPU (frameino_b200/vae.py RowParallel): band split,
halo exchange and band gather, driven through a torch stand-in of the decoder's conv -> up-sample -> conv chain and
compared with the un-sharded chain (the CUDA convolutions themselves are covered by -m gpu tests and
tools/vae_sp_check.py on real GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _conv_cl(x, w):  # x [t, H, W, C] channels-last, zero padding 1: the un-sharded convolution
    return F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1).contiguous()


def _conv_band(frames, w):  # frames [t, h_loc + 2, W, C] with halo rows: "valid" along H, zero padding along W
    return F.conv2d(F.pad(frames.permute(0, 3, 1, 2), (1, 1, 0, 0)), w).permute(0, 2, 3, 1).contiguous()


def _worker(rank, world, port, h, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frameino_b200.vae import RowParallel

        rp = RowParallel()
        g = torch.Generator().manual_seed(0)
        t, wd, c = 2, 6, 4
        x = torch.randn(t, h, wd, c, generator=g)
        w1 = torch.randn(c, c, 3, 3, generator=g) * 0.2
        w2 = torch.randn(c, c, 3, 3, generator=g) * 0.2
        ref = _conv_cl(x, w1)
        ref_up = ref.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        ref2 = _conv_cl(ref_up, w2)
        a, b = rp.rows(h)
        hl = b - a
        # conv 1 on the band: interior written, halos exchanged
        buf = torch.zeros(t, hl + 2, wd, c)
        buf[:, 1:1 + hl] = x[:, a:b]
        rp.exchange(buf)
        y = _conv_band(buf, w1)
        e1 = float((y - ref[:, a:b]).abs().max())
        # 2x up-sampling doubles the band; conv 2 on it
        up = y.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        buf2 = torch.zeros(t, 2 * hl + 2, 2 * wd, c)
        buf2[:, 1:1 + 2 * hl] = up
        rp.exchange(buf2)
        y2 = _conv_band(buf2, w2)
        e2 = float((y2 - ref2[:, 2 * a:2 * b]).abs().max())
        full = rp.gather_rows(y2, h)  # bands are 2x the latent split
        e3 = float((full - ref2).abs().max())
        full5 = rp.gather_rows(y2.permute(3, 0, 1, 2)[None].contiguous(), h, dim=3)  # [1, C, t, rows, W] like the video
        e4 = float((full5[0].permute(1, 2, 3, 0) - ref2).abs().max())
        ret[rank] = (e1, e2, e3, e4, tuple(full.shape), (a, b))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 4), (3, 7), (2, 5)])
def test_row_parallel_conv_chain_matches_unsharded(world, h):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, h, ret), nprocs=world, join=True)
    assert len(ret) == world
    covered = []
    for r in range(world):
        e1, e2, e3, e4, shape, (a, b) = ret[r]
        assert max(e1, e2, e3, e4) <= 1e-5, (r, e1, e2, e3, e4)
        assert shape == (2, 2 * h, 12, 4)
        covered += list(range(a, b))
    assert covered == list(range(h))  # the bands tile the rows exactly once, in rank order


def test_row_split():
    from frameino_b200.vae import RowParallel

    assert RowParallel.split(44, 8) == [(0, 6), (6, 12), (12, 18), (18, 24), (24, 29), (29, 34), (34, 39), (39, 44)]
    assert RowParallel.split(44, 4) == [(0, 11), (11, 22), (22, 33), (33, 44)]
    assert RowParallel.split(4, 4) == [(0, 1), (1, 2), (2, 3), (3, 4)]
    with pytest.raises(ValueError):
        RowParallel.split(3, 4)


def test_row_parallel_falls_back_when_the_canvas_has_fewer_rows_than_ranks():
    from types import SimpleNamespace

    from frameino_b200 import synth
    from frameino_b200.vae import AutoencoderKLWan

    vae = AutoencoderKLWan(**synth.VAE_TINY)
    assert vae._row_parallel_for(44) is None  # not enabled
    vae.row_parallel = SimpleNamespace(world=8)
    assert vae._row_parallel_for(44) is vae.row_parallel
    assert vae._row_parallel_for(8) is vae.row_parallel
    assert vae._row_parallel_for(7) is None  # a band needs at least one latent row: whole frames on every rank
    vae.row_parallel = SimpleNamespace(world=1)
    assert vae._row_parallel_for(44) is None
