"""Ulysses-style sequence parallelism for the Wan denoise step (new capability; the reference has none,
SURVEY.md §8e). One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the exchange.

Every op of the block is token-local except self-attention, so each rank owns a contiguous slice of the token
sequence. Around self-attention two all-to-alls swap the sharded axis:

    [n_loc tokens, all H heads]  --a2a-->  [all N tokens, H/P heads]  --attention-->  --a2a-->  [n_loc, all heads]

RMSNorm-across-heads and RoPE run before the first exchange (they need all heads / are per token). Pad rows (when
N % P != 0) sit at the END of the global sequence, so the attention kernel's ``nk`` bound masks them as keys.

Two transports, same partitioning:
  * ``mode="peer"`` (default on GPUs): both exchanges are FUSED into the neighbouring kernels over NVLink / NVSwitch
    peer memory (``frameino_b200/csrc/peer_kernels.cu``): the q/k norm+RoPE kernel stores every head group directly
    into the owning rank's ``[N, 3*D/P]`` buffer, the attention kernel's epilogue stores every query row directly into
    the owning rank's ``[n_loc, D]`` buffer, and a flag barrier through peer memory separates producer and consumer.
    No NCCL call, no pack/unpack pass, no staging copy.
  * ``mode="nccl"``: ``all_to_all_single`` with the ``fino_swap01`` (de)interleave kernel around it — the baseline the
    fused path is measured against, and the layout the world-size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def exchange_layout(world: int, rank: int, n_loc: int, d_model: int, elem_bytes: int = 2, batch: int = 1) -> dict:
    """Byte layout of one rank's peer buffer and the offsets the fused kernels are given (pure arithmetic; unit-tested
    on CPU). Buffer = [flags 256 B | batch x qkv [world*n_loc, 3*inner] | batch x o [n_loc, d_model]], each part
    256-byte aligned (``batch`` > 1: CogVideoX runs its batched-CFG pair through one exchange, one slot per sample).
    ``o_col_offset`` is where THIS rank's heads start inside every owner's O row."""
    assert d_model % world == 0
    inner = d_model // world

    def up(x):
        return (x + 255) // 256 * 256

    qkv_off = 256
    qkv_bytes = up(world * n_loc * 3 * inner * elem_bytes)
    o_off = up(qkv_off + batch * qkv_bytes)
    o_bytes = up(n_loc * d_model * elem_bytes)
    return {"inner": inner, "n_pad": world * n_loc, "flags_off": 0, "qkv_off": qkv_off, "qkv_row_stride": 3 * inner,
            "o_off": o_off, "o_row_stride": d_model, "o_col_offset": rank * inner * elem_bytes,
            "qkv_batch_bytes": qkv_bytes, "o_batch_bytes": o_bytes, "batch": batch,
            "total_bytes": up(o_off + batch * o_bytes)}


class PeerExchange:
    """The peer-mapped exchange buffers of one (n_loc, d_model, batch) problem: own allocation + the mapped buffers of
    every other rank (CUDA IPC handles swapped once over ``torch.distributed``)."""

    def __init__(self, group, world: int, rank: int, n_loc: int, d_model: int, prims=ops, batch: int = 1):
        # `prims`: the device primitives (peer_alloc/export/import/release/free, pointer_table, tensor_from_ptr,
        # peer_barrier); the world-size-2 gloo test on CPU passes shared-memory stand-ins to exercise this host logic
        self.group, self.world, self.rank, self.n_loc, self.d_model = group, world, rank, n_loc, d_model
        self.batch = batch
        self.prims = ops = prims
        self.layout = lay = exchange_layout(world, rank, n_loc, d_model, batch=batch)
        self.base = ops.peer_alloc(lay["total_bytes"])
        handles = [None] * world
        dist.all_gather_object(handles, ops.peer_export(self.base), group=group)
        self.imported = {}
        bases = []
        for r in range(world):
            if r == rank:
                bases.append(self.base)
            else:
                self.imported[r] = ops.peer_import(handles[r])
                bases.append(self.imported[r])
        self.flag_ptrs = ops.pointer_table([b + lay["flags_off"] for b in bases])
        qb, ob = lay["qkv_batch_bytes"], lay["o_batch_bytes"]
        self.qkv_ptrs_b = [ops.pointer_table([b + lay["qkv_off"] + i * qb for b in bases]) for i in range(batch)]
        self.o_ptrs_b = [ops.pointer_table([b + lay["o_off"] + i * ob + lay["o_col_offset"] for b in bases])
                         for i in range(batch)]
        self.qkv_local_b = [ops.tensor_from_ptr(self.base + lay["qkv_off"] + i * qb, (1, lay["n_pad"], 3 * lay["inner"]))
                            for i in range(batch)]
        self.o_local_b = [ops.tensor_from_ptr(self.base + lay["o_off"] + i * ob, (1, n_loc, d_model))
                          for i in range(batch)]
        # batch slot 0 under the names the single-sample (Wan) path uses
        self.qkv_ptrs, self.o_ptrs = self.qkv_ptrs_b[0], self.o_ptrs_b[0]
        self.qkv_local, self.o_local = self.qkv_local_b[0], self.o_local_b[0]
        self.epoch = 0
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone stores into one

    def barrier(self) -> None:
        self.epoch += 1
        self.prims.peer_barrier(self.flag_ptrs, self.rank, self.world, self.epoch)

    def check(self) -> None:
        """Raises if a peer barrier on this rank timed out (FINO_PEER_TIMEOUT_S): the results since then are invalid.
        Synchronises the stream — called from close() and available to callers at any point they already sync."""
        status = getattr(self.prims, "peer_status", None)
        if status is None or self.base is None:
            return
        bad = [(t, e) for t, e in enumerate(status(self.base + self.layout["flags_off"])) if e]
        if bad:
            raise RuntimeError(f"rank {self.rank}: peer barrier timed out waiting for rank(s) "
                               f"{[t for t, _ in bad]} (epochs {[e for _, e in bad]}); outputs after that are invalid")

    def close(self) -> None:
        if self.base is None:
            return
        ops = self.prims
        if torch.cuda.is_available():
            torch.cuda.synchronize()
            self.check()
        dist.barrier(group=self.group)  # nobody is still storing into a buffer that is about to go away
        self.qkv_local = self.o_local = None
        self.qkv_local_b = self.o_local_b = []
        for p in self.imported.values():
            ops.peer_release(p)
        self.imported = {}
        dist.barrier(group=self.group)
        ops.peer_free(self.base)
        self.base = None


class SequenceParallel:
    def __init__(self, group: Optional[dist.ProcessGroup] = None, mode: str = "peer"):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before enabling sequence parallelism")
        if mode not in ("peer", "nccl"):
            raise ValueError(f"mode {mode!r}: 'peer' (fused NVLink peer-memory exchange) or 'nccl'")
        self.group = group
        self.mode = mode
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._exchanges = {}
        self._prims = ops
        self._qkv_scatter = ops.qkv_norm_rope_scatter
        self._qkv_ln_scatter = ops.qkv_ln_rope_scatter
        self._attention_scatter = ops.attention_scatter
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

    # ---- the same exchange fused into the neighbouring kernels over peer memory -------------------------------
    def exchange(self, n_loc: int, d_model: int, batch: int = 1) -> PeerExchange:
        key = (n_loc, d_model, batch)
        ex = self._exchanges.get(key)
        if ex is None:
            for old in self._exchanges.values():  # one live problem shape at a time (the sampler's canvas is fixed)
                old.close()
            self._exchanges = {}
            ex = self._exchanges[key] = PeerExchange(self.group, self.world, self.rank, n_loc, d_model, self._prims,
                                                     batch=batch)
        return ex

    def fused_attention(self, qkv: torch.Tensor, norm_q_weight, norm_k_weight, heads: int, eps: float, cos, sin,
                        scale: float) -> torch.Tensor:
        """qkv: local RAW fused projections [1, n_loc, 3*D] (before q/k norm). Returns the local attention output
        [1, n_loc, D], a view of this rank's peer buffer (valid until the next call)."""
        assert qkv.dim() == 3 and qkv.shape[0] == 1, "sequence parallel path handles batch 1 (the Wan sampler's case)"
        p = self.world
        n_loc = qkv.shape[1]
        d_model = qkv.shape[2] // 3
        assert heads % p == 0, f"{heads} heads cannot be split over {p} ranks"
        assert n_loc == self.n_loc, "plan() must run before the forward"
        ex = self.exchange(n_loc, d_model)
        lay = ex.layout
        inner = lay["inner"]
        # 1st all-to-all == the stores of the norm+RoPE kernel
        self._qkv_scatter(qkv, norm_q_weight, norm_k_weight, heads, eps, cos, sin, ex.qkv_ptrs, p, self.rank, n_loc,
                          lay["qkv_row_stride"])
        ex.barrier()
        full = ex.qkv_local[:, : self.n_total]
        # 2nd all-to-all == the stores of the attention epilogue
        self._attention_scatter(full[..., :inner], full[..., inner:2 * inner], full[..., 2 * inner:], heads // p,
                                ex.o_ptrs, p, n_loc, lay["o_row_stride"], scale)
        ex.barrier()
        return ex.o_local

    def fused_attention_ln(self, qkv: torch.Tensor, norm_q, norm_k, heads: int, eps: float, cos, sin, rope_skip: int,
                           scale: float) -> torch.Tensor:
        """CogVideoX form of ``fused_attention``: qkv = local RAW fused projections [B, n_loc, 3*D] of the JOINT
        [text | video] rows (B = the pipeline's batched CFG pair); per-head LayerNorm(64) + RoPE (local rows >=
        ``rope_skip``, the text rows, are not rotated) + first exchange in one kernel per sample, one barrier, attention
        + second exchange per sample, one barrier. Returns [B, n_loc, D] (a view of this rank's peer buffer)."""
        assert qkv.dim() == 3
        p = self.world
        b, n_loc, d3 = qkv.shape
        d_model = d3 // 3
        assert heads % p == 0, f"{heads} heads cannot be split over {p} ranks"
        assert n_loc == self.n_loc, "plan() must run before the forward"
        ex = self.exchange(n_loc, d_model, batch=b)
        lay = ex.layout
        inner = lay["inner"]
        for i in range(b):
            self._qkv_ln_scatter(qkv[i], norm_q.weight, norm_q.bias, norm_k.weight, norm_k.bias, heads, eps, cos, sin,
                                 rope_skip, ex.qkv_ptrs_b[i], p, self.rank, n_loc, lay["qkv_row_stride"])
        ex.barrier()
        for i in range(b):
            full = ex.qkv_local_b[i][:, : self.n_total]
            self._attention_scatter(full[..., :inner], full[..., inner:2 * inner], full[..., 2 * inner:], heads // p,
                                    ex.o_ptrs_b[i], p, n_loc, lay["o_row_stride"], scale)
        ex.barrier()
        if b == 1:
            return ex.o_local_b[0]
        # the batch slots are 256-byte aligned, hence not one strided tensor in general: stack them (2 x n_loc x D bf16)
        return torch.cat(ex.o_local_b[:b], dim=0)

    def close(self) -> None:
        for ex in self._exchanges.values():
            ex.close()
        self._exchanges = {}

    def gather_rows(self, y: torch.Tensor) -> torch.Tensor:
        """[1, n_loc, C] local rows -> [1, N, C] on every rank."""
        parts = torch.empty(self.world, y.shape[1], y.shape[2], dtype=y.dtype, device=y.device)
        dist.all_gather_into_tensor(parts.view(-1), y.contiguous().view(-1), group=self.group)
        return parts.view(1, self.world * y.shape[1], y.shape[2])[:, : self.n_total]


def _self_attention_modules(model):
    """The self-attention containers of a native model: Wan ``blocks[i].attn1``, CogVideoX ``transformer_blocks[i].attn1``."""
    blocks = getattr(model, "blocks", None)
    if blocks is None:
        blocks = model.transformer_blocks
    return [blk.attn1 for blk in blocks]


def enable_sequence_parallel(model, group: Optional[dist.ProcessGroup] = None, mode: str = "peer") -> SequenceParallel:
    """Switches a ``frameino_b200`` transformer to Ulysses sequence parallelism over ``group``: ``mode="peer"`` (the
    exchange fused into the neighbouring kernels over NVLink peer memory: RMSNorm + Wan RoPE prologue for Wan, per-head
    LayerNorm + CogVideoX RoPE prologue over the joint text + video rows for CogVideoX) or ``"nccl"``."""
    sp = SequenceParallel(group, mode)
    model.sequence_parallel = sp
    for attn in _self_attention_modules(model):
        attn.__dict__["_fino_sp"] = sp
    return sp


def disable_sequence_parallel(model) -> None:
    if model.sequence_parallel is not None:
        model.sequence_parallel.close()
    model.sequence_parallel = None
    for attn in _self_attention_modules(model):
        attn.__dict__.pop("_fino_sp", None)


def shard_joint_rope(cos: torch.Tensor, sin: torch.Tensor, text_len: int, n_loc: int, rank: int):
    """CogVideoX under sequence parallelism: the joint sequence is [text_len text rows | video rows] and rank r owns
    joint rows [r*n_loc, (r+1)*n_loc). Returns (local_text_len, cos_local, sin_local): the number of leading local rows
    that are text (not rotated; all text must sit on rank 0) and the RoPE table rows of the local video tokens,
    zero-padded to n_loc - local_text_len rows (pad rows past the end of the sequence are never used as keys)."""
    if text_len > n_loc:
        raise NotImplementedError(f"{text_len} text tokens do not fit the first rank's {n_loc} rows")
    local_text = text_len if rank == 0 else 0
    first = rank * n_loc + local_text - text_len  # first video row of this rank
    want = n_loc - local_text
    out = []
    for t in (cos, sin):
        piece = t[first:first + want]
        if piece.shape[0] < want:
            piece = torch.cat([piece, piece.new_zeros(want - piece.shape[0], t.shape[1])], dim=0)
        out.append(piece.contiguous())
    return local_text, out[0], out[1]


class CfgParallel:
    """Classifier-free-guidance parallelism for the sampler loop (SURVEY.md 8e "other axes", 8f row 1): the reference
    pipeline runs the conditional and the unconditional forward of a scheduler step one after the other on the same
    inputs (pipeline_wan_i2v_motion_FrameINO.py:862-882); they are independent, so ranks [0, W/2) run the conditional
    one and ranks [W/2, W) the unconditional one, each half sharding its forward over its own Ulysses group, and rank r
    swaps its output rows with rank r + W/2 (one 2-rank all-gather of [N, 192] bf16 per step). A scheduler step then
    costs one forward at W/2-way sequence parallelism instead of two at W-way, which scales better (fewer, larger
    shards; half the exchange partners)."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before enabling CFG parallelism")
        world = dist.get_world_size(group)
        if world % 2 != 0:
            raise ValueError(f"CFG parallelism needs an even number of ranks, got {world}")
        rank = dist.get_rank(group)
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        half = world // 2
        self.world, self.rank, self.half = world, rank, half
        self.branch = 0 if rank < half else 1  # 0: conditional forward, 1: unconditional
        # every rank creates every group, in the same order (torch.distributed requirement)
        halves = [dist.new_group(ranks[:half]), dist.new_group(ranks[half:])]
        pairs = [dist.new_group([ranks[i], ranks[i + half]]) for i in range(half)]
        self.half_group = halves[self.branch]
        self.pair_group = pairs[rank % half]

    def exchange(self, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """This rank's forward output -> (conditional output, unconditional output)."""
        out = [torch.empty_like(y), torch.empty_like(y)]
        dist.all_gather(out, y.contiguous(), group=self.pair_group)  # group rank 0 is the conditional half's member
        return out[0], out[1]


def enable_cfg_parallel(model, group: Optional[dist.ProcessGroup] = None, mode: str = "peer") -> CfgParallel:
    """Splits ``group`` (default: all ranks) into a conditional and an unconditional half for the sampler loops of
    ``frameino_b200.sampling`` and turns on Ulysses sequence parallelism inside each half when it has more than one
    rank. Pass the returned object to the loop as ``cfg_parallel=``."""
    cp = CfgParallel(group)
    if cp.half > 1:
        enable_sequence_parallel(model, group=cp.half_group, mode=mode)
    model.__dict__["cfg_parallel"] = cp
    return cp


def disable_cfg_parallel(model) -> None:
    """Undoes ``enable_cfg_parallel`` (collective: every rank must call it)."""
    disable_sequence_parallel(model)
    model.__dict__.pop("cfg_parallel", None)
