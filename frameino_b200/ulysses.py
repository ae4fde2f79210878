"""Ulysses-style sequence parallelism for the Wan denoise step (new capability; the reference has none,
SURVEY.md §8e). One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the exchange.

Every op of the block is token-local except self-attention, so each rank owns a contiguous slice of the token
sequence. Around self-attention two all-to-alls swap the sharded axis:

    [n_loc tokens, all H heads]  --a2a-->  [all N tokens, H/P heads]  --attention-->  --a2a-->  [n_loc, all heads]

RMSNorm-across-heads and RoPE run before the first exchange (they need all heads / are per token). Pad rows (when
N % P != 0) sit at the END of the global sequence, so the attention kernel's ``nk`` bound masks them as keys.
The (de)interleave around ``all_to_all_single`` is the ``fino_swap01`` kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


class SequenceParallel:
    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before enabling sequence parallelism")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_total = 0
        self.n_pad = 0
        self.n_loc = 0
        # device primitives; the world_size-2 gloo tests on CPU substitute torch stand-ins to exercise the layout math
        self._swap01 = ops.swap01
        self._attention = ops.attention

    # ---- partitioning (host logic, unit-tested on CPU with gloo) ------------------------------------------------
    @staticmethod
    def partition(n_total: int, world: int) -> Tuple[int, int]:
        """(n_loc, n_pad): tokens per rank and padded global length."""
        n_loc = (n_total + world - 1) // world
        return n_loc, n_loc * world

    def plan(self, n_total: int) -> None:
        self.n_total = n_total
        self.n_loc, self.n_pad = self.partition(n_total, self.world)

    def local_slice(self) -> slice:
        return slice(self.rank * self.n_loc, min((self.rank + 1) * self.n_loc, self.n_total))

    def shard_rows(self, x: torch.Tensor, dim: int = 1) -> torch.Tensor:
        """Local slice of a [.., N, ..] tensor along ``dim``, zero-padded to n_loc rows."""
        sl = self.local_slice()
        part = x.narrow(dim, sl.start, max(sl.stop - sl.start, 0))
        if part.shape[dim] < self.n_loc:
            pad_shape = list(part.shape)
            pad_shape[dim] = self.n_loc - part.shape[dim]
            part = torch.cat([part, part.new_zeros(pad_shape)], dim=dim)
        return part.contiguous()

    def rows_per_group(self, rows_per_group: int) -> int:
        # scalar-timestep path (one modulation row per batch element): the group is the local sequence
        return self.n_loc if rows_per_group else 0

    # ---- the exchange around self-attention ---------------------------------------------------------------------
    def attention(self, qkv: torch.Tensor, heads: int, scale: float) -> torch.Tensor:
        """qkv: local fused projections [1, n_loc, 3*D] (q/k already normalised + rotated). Returns the local
        attention output [1, n_loc, D]."""
        assert qkv.dim() == 3 and qkv.shape[0] == 1, "sequence parallel path handles batch 1 (the Wan sampler's case)"
        p = self.world
        n_loc = qkv.shape[1]
        d_model = qkv.shape[2] // 3
        assert heads % p == 0, f"{heads} heads cannot be split over {p} ranks"
        hp = heads // p
        hd = d_model // heads
        inner = hp * hd
        # [n_loc, 3, P, inner] -> [P, n_loc, 3, inner]: destination-major send buffer
        send = self._swap01(qkv.reshape(-1), n_loc * 3, p, inner)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.reshape(-1), group=self.group)
        full = recv.view(1, p * n_loc, 3 * inner)  # source-major == global token order
        q, k, v = full[..., :inner], full[..., inner:2 * inner], full[..., 2 * inner:]
        n_valid = self.n_total
        o = self._attention(q[:, :n_valid] if n_valid < p * n_loc else q, k[:, :n_valid], v[:, :n_valid], hp,
                            scale=scale)
        if n_valid < p * n_loc:
            o_full = torch.zeros(1, p * n_loc, inner, dtype=o.dtype, device=o.device)
            o_full[:, :n_valid] = o
            o = o_full
        # rows of rank j are contiguous in o: send them back, receive [P(src heads), n_loc, inner]
        back = torch.empty_like(o)
        dist.all_to_all_single(back.view(-1), o.contiguous().view(-1), group=self.group)
        out = self._swap01(back.view(-1), p, n_loc, inner)  # -> [n_loc, P, inner] == [n_loc, D]
        return out.view(1, n_loc, d_model)

    def gather_rows(self, y: torch.Tensor) -> torch.Tensor:
        """[1, n_loc, C] local rows -> [1, N, C] on every rank."""
        parts = torch.empty(self.world, y.shape[1], y.shape[2], dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(parts.view(-1), y.contiguous().view(-1), group=self.group)
        return parts.view(1, self.world * y.shape[1], y.shape[2])[:, : self.n_total]


def enable_sequence_parallel(model, group: Optional[dist.ProcessGroup] = None) -> SequenceParallel:
    """Switches a ``frameino_b200.WanTransformer3DModel`` to Ulysses sequence parallelism over ``group``."""
    sp = SequenceParallel(group)
    model.sequence_parallel = sp
    for blk in model.blocks:
        blk.attn1.__dict__["_fino_sp"] = sp
    return sp


def disable_sequence_parallel(model) -> None:
    model.sequence_parallel = None
    for blk in model.blocks:
        blk.attn1.__dict__.pop("_fino_sp", None)
