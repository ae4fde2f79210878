"""CFG parallelism of the sampler loop (frameino_b200.ulysses.CfgParallel) on CPU with gloo: group construction and
the pair exchange at world sizes 2 and 4, and the plain sampler loop with one CFG branch per rank against the serial
loop (CPU oracle transformer, tiny config): identical final latents."""
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


def _loop_case():
    from frameino_b200 import synth
    from oracle import wan_oracle

    cfg = synth.WAN_TINY
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.float32)
    g = torch.Generator().manual_seed(7)
    c, f, h, w = 16, 2, 8, 8
    lat = torch.randn(1, c, f, h, w, generator=g)
    cond = torch.zeros(1, c, f, h, w)
    cond[:, :, 0] = torch.randn(1, c, h, w, generator=g)
    mask = torch.ones(1, 1, f, h, w)
    mask[:, :, 0] = 0
    traj = torch.randn(1, c, f + 1, h, w, generator=g)
    traj[:, :, f:] = 0
    idl = torch.randn(1, c, 1, h, w, generator=g)
    pos = torch.randn(1, 8, 64, generator=g)
    neg = torch.zeros(1, 8, 64)
    ocfg = wan_oracle.WanConfig(**cfg)

    def oracle_tf(hidden_states, timestep, encoder_hidden_states, return_dict=False):
        return (wan_oracle.wan_forward(sd, ocfg, hidden_states, timestep, encoder_hidden_states),)

    return oracle_tf, (lat, cond, mask, traj, idl, pos, neg)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from frameino_b200.sampling import wan_frameino_denoise
        from frameino_b200.ulysses import CfgParallel

        cp = CfgParallel()
        half = world // 2
        assert cp.branch == (0 if rank < half else 1) and cp.half == half
        assert dist.get_world_size(cp.half_group) == half and dist.get_world_size(cp.pair_group) == 2
        # pair exchange: rank r <-> r + half, conditional member first
        mine = torch.full((3, 5), float(rank))
        a, b = cp.exchange(mine)
        lo = rank % half
        assert torch.equal(a, torch.full((3, 5), float(lo))) and torch.equal(b, torch.full((3, 5), float(lo + half)))
        # the half group really is this rank's half
        t = torch.tensor([float(rank)])
        dist.all_reduce(t, group=cp.half_group)
        want = sum(range(half)) if cp.branch == 0 else sum(range(half, world))
        assert float(t) == float(want)
        out = None
        if world == 2:
            tf, tensors = _loop_case()
            out = wan_frameino_denoise(tf, *tensors, num_steps=3, model_dtype=torch.float32, cfg_parallel=cp)
            with pytest.raises(ValueError, match="classifier-free"):
                wan_frameino_denoise(tf, *tensors[:6], None, num_steps=1, model_dtype=torch.float32, cfg_parallel=cp)
        ret[rank] = ("ok", out)
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = ("error", traceback.format_exc() + repr(e))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cfg_parallel_groups_exchange_and_loop(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r][0] == "ok", ret[r][1]
    if world == 2:
        from frameino_b200.sampling import wan_frameino_denoise

        tf, tensors = _loop_case()
        serial = wan_frameino_denoise(tf, *tensors, num_steps=3, model_dtype=torch.float32)
        for r in range(world):
            assert torch.equal(ret[r][1], serial), f"rank {r}: CFG-parallel loop differs from the serial loop"


def test_cfg_parallel_needs_an_even_world():
    port = _free_port()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from frameino_b200.ulysses import CfgParallel

        with pytest.raises(ValueError, match="even"):
            CfgParallel()
    finally:
        dist.destroy_process_group()
