"""B200-native ``WanTransformer3DModel`` — drop-in for the reference class of the same name
(reference architecture/transformer_wan.py:353-552): same constructor config, same ``forward`` signature and return
type, same diffusers state-dict key names, same attributes the FrameINO pipeline reads (``.config``, ``.dtype``,
``.cache_context``). The forward issues only frameino_b200 CUDA kernels.

What changes relative to the reference arithmetic (and why it is still the same function):
  * the time MLP runs on the UNIQUE timestep values and tokens index the result (the reference recomputes it per
    token, transformer_wan.py:175-183, :317-319; bit-identical per row, removes ~3.8 TFLOP and ~60 GB of HBM traffic);
  * q/k/v projections run as one GEMM over a concatenated weight; LayerNorm+modulate, RMSNorm+RoPE and the gated
    residuals are fused kernels / GEMM epilogues with the reference's cast points preserved.
"""
from __future__ import annotations

import math
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import ops
from .modules import (Attention, FeedForward, ModelBase, TextProjection, TimestepEmbedding, WeightOnlyNorm,
                      compute_dtype, logger)
from .processors import FinoWanAttnProcessor, tensor_key


@dataclass
class Transformer2DModelOutput:
    """diffusers.models.modeling_outputs.Transformer2DModelOutput stand-in (attribute + index access)."""

    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


@dataclass
class WanTextState:
    """The text-side work of one forward that does not depend on the latents or the timestep — constant over the 50
    sampler steps of a prompt: the text-embedder MLP (transformer_wan.py:185) and every block's
    cross-attention ``norm_k(to_k(text))`` / ``to_v(text)`` (transformer_wan.py:61-67 with encoder_hidden_states = text).
    Built by ``WanTransformer3DModel.prepare_text``; reused across forwards inside ``model.cache_context(name)``."""

    text: torch.Tensor                                         # [B, T, D] bf16
    kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]]      # per block; None where a foreign processor is plugged in
    key: Any = None                                            # validity key (input + weight identities/versions)
    source: Optional[torch.Tensor] = None                      # strong reference: keeps the keyed memory from being reused
    ready: Optional[List[Any]] = None                          # per block: CUDA event recorded on the side stream that
                                                               # produced kv[i]; None once the main stream has waited


def _rope_1d(dim: int, max_len: int, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    # embeddings.py:1153-1207 (use_real, repeat_interleave_real, float64 frequencies)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
    ang = torch.outer(torch.arange(max_len, dtype=torch.float64), freqs)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


class WanRotaryPosEmbed(nn.Module):
    """3-D RoPE tables (transformer_wan.py:192-253). ``forward`` returns ``(cos, sin)`` each [1, 1, N, head_dim]
    fp32 in (frame, height, width) token order; the gathered table is cached per latent shape."""

    def __init__(self, attention_head_dim: int, patch_size: Tuple[int, int, int], max_seq_len: int,
                 theta: float = 10000.0):
        super().__init__()
        self.attention_head_dim = attention_head_dim
        self.patch_size = tuple(patch_size)
        self.max_seq_len = max_seq_len
        h_dim = w_dim = 2 * (attention_head_dim // 6)
        t_dim = attention_head_dim - h_dim - w_dim
        tabs = [_rope_1d(d, max_seq_len, theta) for d in (t_dim, h_dim, w_dim)]
        self.register_buffer("freqs_cos", torch.cat([t[0] for t in tabs], dim=1), persistent=False)
        self.register_buffer("freqs_sin", torch.cat([t[1] for t in tabs], dim=1), persistent=False)
        self._cache: Dict[Any, Tuple[torch.Tensor, torch.Tensor]] = {}

    def forward(self, hidden_states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        _, _, f, h, w = hidden_states.shape
        return self.tables(f, h, w)

    def tables(self, f: int, h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(cos, sin) for a latent of f x h x w (before patching)."""
        key = (f, h, w, self.freqs_cos.device)
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        p_t, p_h, p_w = self.patch_size
        ppf, pph, ppw = f // p_t, h // p_h, w // p_w
        d = self.attention_head_dim
        split = [d - 2 * (d // 3), d // 3, d // 3]
        if split[1] != 2 * (d // 6):
            raise ValueError(f"attention_head_dim={d}: RoPE band split {split} disagrees with the table construction "
                             "(the reference has the same restriction, transformer_wan.py:206-207 vs :233-237)")
        out = []
        for tab in (self.freqs_cos, self.freqs_sin):
            tf, th, tw = tab.split(split, dim=1)
            tf = tf[:ppf].view(ppf, 1, 1, -1).expand(ppf, pph, ppw, -1)
            th = th[:pph].view(1, pph, 1, -1).expand(ppf, pph, ppw, -1)
            tw = tw[:ppw].view(1, 1, ppw, -1).expand(ppf, pph, ppw, -1)
            out.append(torch.cat([tf, th, tw], dim=-1).reshape(1, 1, ppf * pph * ppw, -1).contiguous())
        self._cache = {key: (out[0], out[1])}
        return out[0], out[1]


class WanTimeTextImageEmbedding(nn.Module):
    """Parameter layout of transformer_wan.py:146-189 (time_embedder, time_proj, text_embedder)."""

    def __init__(self, dim: int, time_freq_dim: int, time_proj_dim: int, text_embed_dim: int,
                 image_embed_dim: Optional[int] = None):
        super().__init__()
        if image_embed_dim is not None:
            raise NotImplementedError("image_dim (Wan2.1 I2V CLIP branch) is not part of the FrameINO Wan2.2 path")
        self.time_freq_dim = time_freq_dim
        self.time_embedder = TimestepEmbedding(time_freq_dim, dim)
        self.time_proj = nn.Linear(dim, time_proj_dim)
        self.text_embedder = TextProjection(text_embed_dim, dim)
        self.image_embedder = None


class WanTransformerBlock(nn.Module):
    """transformer_wan.py:257-350. Parameters only differ from the reference in that norm layers without affine
    parameters hold nothing; ``forward`` launches the fused kernels."""

    def __init__(self, dim: int, ffn_dim: int, num_heads: int, qk_norm: str = "rms_norm_across_heads",
                 cross_attn_norm: bool = False, eps: float = 1e-6, added_kv_proj_dim: Optional[int] = None):
        super().__init__()
        if added_kv_proj_dim is not None:
            raise NotImplementedError("added_kv_proj_dim (Wan2.1 I2V) is not part of the FrameINO Wan2.2 path")
        self.eps = eps
        self.norm1 = WeightOnlyNorm(dim, eps, elementwise_affine=False)
        self.attn1 = Attention(dim, num_heads, dim // num_heads, qk_norm, eps, processor=FinoWanAttnProcessor())
        self.attn2 = Attention(dim, num_heads, dim // num_heads, qk_norm, eps, processor=FinoWanAttnProcessor())
        self.norm2 = WeightOnlyNorm(dim, eps, elementwise_affine=True, bias=True) if cross_attn_norm else nn.Identity()
        self.ffn = FeedForward(dim, ffn_dim)
        self.norm3 = WeightOnlyNorm(dim, eps, elementwise_affine=False)
        self.scale_shift_table = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def _norm2_affine(self):
        """norm2's gain / bias as fp32 (FP32LayerNorm computes in fp32 whatever dtype the parameters were cast to:
        ``_keep_in_fp32_modules`` keeps them fp32, a plain ``model.to(torch.bfloat16)`` does not)."""
        w, b = self.norm2.weight, self.norm2.bias
        if w.dtype == torch.float32 and b.dtype == torch.float32:
            return w, b
        key = (tensor_key(w), tensor_key(b))
        hit = self.__dict__.get("_fino_cache")
        if hit is None or hit[0] != key:
            hit = self.__dict__["_fino_cache"] = (key, w.detach().float().contiguous(), b.detach().float().contiguous())
        return hit[1], hit[2]

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, mod: torch.Tensor,
                row_index: Optional[torch.Tensor], rows_per_group: int, rotary_emb, text_kv=None) -> torch.Tensor:
        """``mod``: fp32 [R, 6*dim] = scale_shift_table + timestep_proj rows (shift, scale, gate, c_shift, c_scale,
        c_gate). ``hidden_states`` [B, N, dim] is updated in place and returned. ``text_kv``: this block's
        pre-projected cross-attention (k, v) from a ``WanTextState`` (native processor only)."""
        x = hidden_states
        dim = x.shape[-1]
        tap = self.__dict__.pop("_fino_tap", None)  # parity tests: (dict, "blocks.i") collects intermediate tensors
        shift, scale, gate = mod[:, 0:dim], mod[:, dim:2 * dim], mod[:, 2 * dim:3 * dim]
        c_shift, c_scale, c_gate = mod[:, 3 * dim:4 * dim], mod[:, 4 * dim:5 * dim], mod[:, 5 * dim:6 * dim]
        sel = dict(row_index=row_index, rows_per_group=rows_per_group)

        # 1. self-attention (:334-336)
        h = ops.ln_modulate(x, self.eps, shift=shift, scale=scale, **sel)
        if tap is not None:
            tap[0][tap[1] + ".norm1"] = h.clone()
        if isinstance(self.attn1.processor, FinoWanAttnProcessor):
            x = self.attn1(hidden_states=h, rotary_emb=rotary_emb, fino_residual=(x, gate, row_index, rows_per_group))
        else:  # foreign processor plugged in through set_processor: keep the reference dataflow
            a = self.attn1(hidden_states=h, rotary_emb=rotary_emb)
            x = ops.gate_residual(x, a.contiguous(), gate, out=x, **sel)

        if tap is not None:
            tap[0][tap[1] + ".after_attn1"] = x.clone()

        # 2. cross-attention (:339-341)
        if isinstance(self.norm2, WeightOnlyNorm):
            g2, b2 = self._norm2_affine()
            h = ops.ln_modulate(x, self.eps, gamma=g2, beta=b2, out=h)
        else:
            h = x
        if isinstance(self.attn2.processor, FinoWanAttnProcessor):
            x = self.attn2(hidden_states=h, encoder_hidden_states=encoder_hidden_states,
                           fino_residual=(x, None, None, 0), fino_text_kv=text_kv)
        else:
            a = self.attn2(hidden_states=h, encoder_hidden_states=encoder_hidden_states)
            x = ops.gate_residual(x, a.contiguous(), out=x)

        if tap is not None:
            tap[0][tap[1] + ".after_attn2"] = x.clone()

        # 3. feed-forward (:344-348)
        h = ops.ln_modulate(x, self.eps, shift=c_shift, scale=c_scale, out=h if h is not x else None, **sel)
        up, down = self.ffn.net[0].proj, self.ffn.net[2]
        f = ops.linear(h, up.weight, up.bias, epilogue=ops.EPI_GELU_TANH)
        x = ops.linear(f, down.weight, down.bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=c_gate, out=x, **sel)
        return x


class WanTransformer3DModel(ModelBase):
    """Drop-in for reference ``architecture.transformer_wan.WanTransformer3DModel`` (see module docstring)."""

    _supports_gradient_checkpointing = False
    _skip_layerwise_casting_patterns = ["patch_embedding", "condition_embedder", "norm"]
    _no_split_modules = ["WanTransformerBlock"]
    _keep_in_fp32_modules = ["time_embedder", "scale_shift_table", "norm1", "norm2", "norm3"]
    _keys_to_ignore_on_load_unexpected = ["norm_added_q"]
    _repeated_blocks = ["WanTransformerBlock"]

    def __init__(
        self,
        patch_size: Tuple[int, int, int] = (1, 2, 2),
        num_attention_heads: int = 40,
        attention_head_dim: int = 128,
        in_channels: int = 16,
        out_channels: int = 16,
        text_dim: int = 4096,
        freq_dim: int = 256,
        ffn_dim: int = 13824,
        num_layers: int = 40,
        cross_attn_norm: bool = True,
        qk_norm: Optional[str] = "rms_norm_across_heads",
        eps: float = 1e-6,
        image_dim: Optional[int] = None,
        added_kv_proj_dim: Optional[int] = None,
        rope_max_seq_len: int = 1024,
        pos_embed_seq_len: Optional[int] = None,
    ) -> None:
        super().__init__()
        self._register_config(
            patch_size=tuple(patch_size), num_attention_heads=num_attention_heads,
            attention_head_dim=attention_head_dim, in_channels=in_channels, out_channels=out_channels,
            text_dim=text_dim, freq_dim=freq_dim, ffn_dim=ffn_dim, num_layers=num_layers,
            cross_attn_norm=cross_attn_norm, qk_norm=qk_norm, eps=eps, image_dim=image_dim,
            added_kv_proj_dim=added_kv_proj_dim, rope_max_seq_len=rope_max_seq_len, pos_embed_seq_len=pos_embed_seq_len,
        )
        inner_dim = num_attention_heads * attention_head_dim
        out_channels = out_channels or in_channels
        self.rope = WanRotaryPosEmbed(attention_head_dim, patch_size, rope_max_seq_len)
        self.patch_embedding = nn.Conv3d(in_channels, inner_dim, kernel_size=patch_size, stride=patch_size)
        self.condition_embedder = WanTimeTextImageEmbedding(inner_dim, freq_dim, inner_dim * 6, text_dim, image_dim)
        self.blocks = nn.ModuleList(
            [WanTransformerBlock(inner_dim, ffn_dim, num_attention_heads, qk_norm, cross_attn_norm, eps,
                                 added_kv_proj_dim) for _ in range(num_layers)]
        )
        self.norm_out = WeightOnlyNorm(inner_dim, eps, elementwise_affine=False)
        self.proj_out = nn.Linear(inner_dim, out_channels * math.prod(patch_size))
        self.scale_shift_table = nn.Parameter(torch.randn(1, 2, inner_dim) / inner_dim ** 0.5)
        self.gradient_checkpointing = False
        self._sst_cache = None
        self.sequence_parallel = None  # set by frameino_b200.ulysses.enable_sequence_parallel

    # ------------------------------------------------------------------------------------------------------------
    def to_inference_dtype(self, dtype: torch.dtype = torch.bfloat16) -> "WanTransformer3DModel":
        """Casts like ``from_pretrained(torch_dtype=bf16)`` + ``_keep_in_fp32_modules`` (transformer_wan.py:393).
        ``torch.float16`` (the reference demo's choice, app.py:156) is converted to bf16 with a warning."""
        dtype = compute_dtype(dtype)
        for name, p in self.named_parameters():
            keep = any(k in name for k in self._keep_in_fp32_modules)
            p.data = p.data.to(torch.float32 if keep else dtype)
        return self

    def prepare(self) -> "WanTransformer3DModel":
        """Concatenated q|k|v (self-attention) and k|v (cross-attention) projection weights + the stacked modulation
        tables, built once after loading instead of on the first forward."""
        from .processors import _fused_weights

        for b in self.blocks:
            _fused_weights(b.attn1, ("to_q", "to_k", "to_v"), "qkv")
            _fused_weights(b.attn2, ("to_k", "to_v"), "kv")
        self._stacked_tables()
        return self

    def _stacked_tables(self) -> torch.Tensor:
        """[L, 6*D] fp32 copy of every block's scale_shift_table (refreshed if a table changes)."""
        key = tuple(tensor_key(b.scale_shift_table) for b in self.blocks)
        if self._sst_cache is None or self._sst_cache[0] != key:
            with torch.no_grad():
                t = torch.stack([b.scale_shift_table.reshape(-1).float() for b in self.blocks]).contiguous()
            self._sst_cache = (key, t)
        return self._sst_cache[1]

    def _conditioning(self, timestep: torch.Tensor, batch: int, tokens: int, dedup: str = "device"):
        """De-duplicated time MLP. Returns (temb_rows fp32-of-bf16 [R, D], proj_rows fp32-of-bf16 [R, 6D],
        row_index int32 [B*N] or None, rows_per_group).

        Per-token timesteps ([B, N], wan 2.2 ti2v, transformer_wan.py:490-492) are de-duplicated ON THE DEVICE
        (``fino_timestep_dedup``: up to 8 distinct values, no host round trip; the FrameINO sampler passes two). The
        number of distinct values is copied to pinned host memory behind an event that ``forward`` checks after it has
        queued the whole step; only if more than 8 were seen does it re-run with ``dedup="unique"`` (torch.unique)."""
        if timestep.ndim == 2:
            if timestep.shape != (batch, tokens):
                raise ValueError(f"timestep shape {tuple(timestep.shape)} != (batch, tokens) = ({batch}, {tokens})")
            flat = timestep.reshape(-1).float().contiguous()
            if dedup == "device":
                uniq, row_index, count = ops.timestep_dedup(flat)
                host = self.__dict__.get("_fino_dedup_host")
                if host is None:
                    host = self.__dict__["_fino_dedup_host"] = torch.zeros(1, dtype=torch.int32).pin_memory()
                    self.__dict__["_fino_dedup_event"] = torch.cuda.Event()
                host.copy_(count, non_blocking=True)
                self.__dict__["_fino_dedup_event"].record()
                self.__dict__["_fino_dedup_pending"] = True
            else:
                uniq, inverse = torch.unique(flat, return_inverse=True)
                row_index = inverse.to(torch.int32).contiguous()
            rows_per_group = 0
        else:
            if timestep.ndim == 0:
                timestep = timestep.reshape(1).expand(batch)
            uniq = timestep.reshape(-1).float().contiguous()
            if uniq.numel() != batch:
                raise ValueError(f"timestep has {uniq.numel()} entries for batch {batch}")
            row_index = None
            rows_per_group = tokens
        temb, proj = self.time_rows(uniq)
        return temb, proj, row_index, rows_per_group

    def _dedup_overflowed(self) -> bool:
        """True when the device-side de-duplication of the forward just queued saw more than 8 distinct timesteps.
        Waits only for the event recorded right after the first kernel of that forward, not for the forward."""
        if not self.__dict__.pop("_fino_dedup_pending", False):
            return False
        self.__dict__["_fino_dedup_event"].synchronize()
        return int(self.__dict__["_fino_dedup_host"][0]) > 8

    def time_rows(self, uniq: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Time MLP on distinct timestep values (fp32 [R]): (temb [R, D], timestep_proj [R, 6D]), fp32 holding the
        bf16-rounded module outputs (transformer_wan.py:175-183). Row r is what the reference computes for every token
        whose timestep is uniq[r]: the fp32 ``time_embedder`` (``_keep_in_fp32_modules``, :393) runs in fp32 with fp32
        accumulation for ANY number of rows (8-row chunks inside one launch), its result is rounded to the model dtype
        (:182) and ``time_proj`` runs as a bf16 module (:183)."""
        ce = self.condition_embedder
        te = ce.time_embedder
        emb = ops.timestep_embedding(uniq.contiguous(), ce.time_freq_dim, True, 0.0)  # :175
        dt = self.proj_out.weight.dtype
        w_f32 = te.linear_1.weight.dtype == torch.float32
        h = ops.linear_small_m(emb, te.linear_1.weight, te.linear_1.bias, act_out=1, round_in=not w_f32,
                               round_out=not w_f32)
        temb = ops.linear_small_m(h, te.linear_2.weight, te.linear_2.bias, round_out=not w_f32)
        temb = temb.to(dt).float()  # .type_as(encoder_hidden_states), :182
        proj = ops.linear_small_m(temb, ce.time_proj.weight, ce.time_proj.bias, act_in=1, round_in=True,
                                  round_out=True)  # :183 (bf16 module)
        return temb, proj

    # ------------------------------------------------------------------------------------------------------------
    # step-invariant text work
    @contextmanager
    def cache_context(self, name: str):
        """The reference pipeline wraps its two CFG forwards in ``cache_context("cond")`` / ``("uncond")``
        (pipeline_wan_i2v_motion_FrameINO.py:862, :873). Here the name keys a ``WanTextState``: the text-embedder MLP
        and every block's cross-attention K/V are computed on the first forward of a prompt and reused by the later
        steps, as long as the same ``encoder_hidden_states`` memory (unchanged) and the same weights are passed."""
        prev = self.__dict__.get("_fino_cache_name")
        self.__dict__["_fino_cache_name"] = name
        try:
            yield
        finally:
            self.__dict__["_fino_cache_name"] = prev

    def clear_text_cache(self) -> None:
        self.__dict__["_fino_text_cache"] = {}

    def _text_key(self, ehs: torch.Tensor):
        ident = tensor_key  # works for inference-mode prompt embeddings too (no version counter: keyed by address)
        te = self.condition_embedder.text_embedder
        parts = [ident(ehs), tuple(ehs.stride()), ehs.device,
                 ident(te.linear_1.weight), ident(te.linear_1.bias), ident(te.linear_2.weight), ident(te.linear_2.bias)]
        for b in self.blocks:
            a = b.attn2
            parts.append((id(a.processor), ident(a.to_k.weight), ident(a.to_k.bias), ident(a.to_v.weight),
                          ident(a.to_v.bias), ident(getattr(a.norm_k, "weight", None))))
        return tuple(parts)

    def prepare_text(self, encoder_hidden_states: torch.Tensor) -> WanTextState:
        """Runs the text embedder (:185) and every block's cross-attention key/value projection + key norm once.

        The 30 (kv GEMM, k-norm) pairs depend only on the prompt, not on the latents, so they are queued on a SIDE
        stream (``overlap_text_projection``, default on): they fill the SMs whenever the main stream leaves them idle —
        under sequence parallelism that is the NVLink-bound exchange and the skew absorbed by the peer barriers — instead
        of sitting on the critical path in front of every cross-attention. Block i's cross-attention waits on the event
        recorded after kv[i] (``forward_rows``). Nothing is reordered arithmetically: same kernels, same inputs."""
        if not encoder_hidden_states.is_cuda:
            raise RuntimeError("frameino_b200 has no CPU path: move the model and inputs to a CUDA device")
        dt = self.proj_out.weight.dtype
        ce = self.condition_embedder
        text_in = encoder_hidden_states.to(dt)
        t1, t2 = ce.text_embedder.linear_1, ce.text_embedder.linear_2
        text = ops.linear(ops.linear(text_in, t1.weight, t1.bias, epilogue=ops.EPI_GELU_TANH), t2.weight, t2.bias)  # :185
        kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = []
        ready: Optional[List[Any]] = None
        dev = text.device
        if self.__dict__.get("overlap_text_projection", True):
            main = torch.cuda.current_stream(dev)
            side = self.__dict__.get("_fino_side_stream")
            if side is None or side.device != dev:
                side = self.__dict__["_fino_side_stream"] = torch.cuda.Stream(device=dev)
            side.wait_stream(main)  # `text` is ready; everything the previous forward queued on `main` has been
            ready = []              # ordered before, so blocks freed back to the side stream's pool are not in use
            with torch.cuda.stream(side):
                for b in self.blocks:
                    proc = b.attn2.processor
                    kv.append(proc.project_text(b.attn2, text) if isinstance(proc, FinoWanAttnProcessor) else None)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    ready.append(ev)
        else:
            for b in self.blocks:
                proc = b.attn2.processor
                kv.append(proc.project_text(b.attn2, text) if isinstance(proc, FinoWanAttnProcessor) else None)
        return WanTextState(text=text, kv=kv, key=self._text_key(encoder_hidden_states), source=encoder_hidden_states,
                            ready=ready)

    def _text_state(self, encoder_hidden_states: torch.Tensor) -> WanTextState:
        name = self.__dict__.get("_fino_cache_name")
        if name is None:
            return self.prepare_text(encoder_hidden_states)
        cache = self.__dict__.setdefault("_fino_text_cache", {})
        hit = cache.get(name)
        if hit is not None and hit.key == self._text_key(encoder_hidden_states):
            return hit
        state = self.prepare_text(encoder_hidden_states)
        cache[name] = state
        return state

    # ------------------------------------------------------------------------------------------------------------
    def forward_rows(self, rows: torch.Tensor, batch: int, grid: Tuple[int, int, int], conditioning,
                     text_state: WanTextState) -> torch.Tensor:
        """The transformer body on already patchified input: ``rows`` bf16 [batch * tokens, C_in*pt*ph*pw] (token order
        (frame, height, width) of the ``grid`` = latent (frames, height, width)), ``conditioning`` =
        (temb [R, D], timestep_proj [R, 6D], row_index int32 [batch*tokens] | None, rows_per_group) as ``time_rows``
        produces them. Returns the proj_out rows bf16 [batch * tokens, pt*ph*pw*C_out] (:537), i.e. the model output
        before the final un-patchify. ``forward`` = patchify + this + unpatchify; the fused sampler loop
        (frameino_b200/sampling.py) calls it directly."""
        cfg = self.config
        p_t, p_h, p_w = cfg.patch_size
        frames, height, width = grid
        tokens = (frames // p_t) * (height // p_h) * (width // p_w)
        dim = cfg.num_attention_heads * cfg.attention_head_dim
        if rows.shape[0] != batch * tokens:
            raise ValueError(f"rows has {rows.shape[0]} rows for batch {batch} x {tokens} tokens")
        temb, proj, row_index, rows_per_group = conditioning
        rotary_emb = self.rope.tables(frames, height, width)  # :484

        sp = self.sequence_parallel
        if sp is not None:  # Ulysses: every rank keeps a contiguous token slice (frameino_b200/ulysses.py)
            if batch != 1:
                raise NotImplementedError("sequence parallelism handles batch 1 (the Wan sampler runs B=1 forwards)")
            sp.plan(tokens)
            rows = sp.shard_rows(rows.view(1, tokens, -1)).view(sp.n_loc, -1)
            if row_index is not None:
                row_index = sp.shard_rows(row_index.view(1, tokens)).view(-1)
            rotary_emb = tuple(sp.shard_rows(t, dim=2) for t in rotary_emb)
            rows_per_group = sp.rows_per_group(rows_per_group)
        n_loc = rows.shape[0] // batch

        pe = self.patch_embedding
        x = ops.linear(rows, pe.weight.view(dim, -1), pe.bias).view(batch, n_loc, dim)  # :486-487
        text = text_state.text

        # per-layer modulation rows: scale_shift_table + timestep_proj.float()  (:317-331)
        mod_all = ops.build_mod_table(self._stacked_tables(), proj.contiguous(), len(self.blocks), 6 * dim)

        taps = self.__dict__.get("_fino_taps")  # parity tests set this to a dict to collect per-layer outputs
        if taps is not None:
            taps["patch_embed"] = x.clone()
            taps["text"] = text.clone()
        ready = text_state.ready
        for i, block in enumerate(self.blocks):  # :516-517
            if taps is not None:
                block.__dict__["_fino_tap"] = (taps, f"blocks.{i}")
            if ready is not None:  # kv[i] was produced on the side stream (prepare_text)
                torch.cuda.current_stream(x.device).wait_event(ready[i])
            x = block(x, text, mod_all[i], row_index, rows_per_group, rotary_emb, text_state.kv[i])
            if taps is not None:
                taps[f"blocks.{i}.out"] = x.clone()

        text_state.ready = None  # the main stream is now ordered after every kv[i]; later forwards need not wait again

        # output modulation (:520-536): scale_shift_table[1,2,D] + temb
        out_mod = ops.build_mod_table(self.scale_shift_table.reshape(1, 2 * dim).float(),
                                      temb.repeat(1, 2).contiguous(), 1, 2 * dim)[0]
        h = ops.ln_modulate(x, cfg.eps, shift=out_mod[:, :dim], scale=out_mod[:, dim:], row_index=row_index,
                            rows_per_group=rows_per_group)
        y = ops.linear(h, self.proj_out.weight, self.proj_out.bias)  # :537
        if sp is not None:
            y = sp.gather_rows(y)
        return y.view(batch * tokens, -1)

    def _check_ready(self, t: torch.Tensor) -> torch.dtype:
        if not t.is_cuda:
            raise RuntimeError("frameino_b200 has no CPU path: move the model and inputs to a CUDA device")
        dt = self.proj_out.weight.dtype
        if dt != torch.bfloat16:
            raise NotImplementedError(f"model dtype {dt}: call .to_inference_dtype(torch.bfloat16) first "
                                      "(torch.float16 is accepted there and converted to bf16)")
        return dt

    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        timestep: torch.Tensor,
        encoder_hidden_states: torch.Tensor,
        encoder_hidden_states_image: Optional[torch.Tensor] = None,
        return_dict: bool = True,
        attention_kwargs: Optional[Dict[str, Any]] = None,
    ) -> Union[Transformer2DModelOutput, Tuple[torch.Tensor]]:
        if attention_kwargs is not None and attention_kwargs.get("scale", None) is not None:
            logger.warning("Passing `scale` via `attention_kwargs` when not using the PEFT backend is ineffective.")
        if encoder_hidden_states_image is not None:
            raise NotImplementedError("encoder_hidden_states_image (Wan2.1 I2V) is not part of the FrameINO path")
        dt = self._check_ready(hidden_states)
        cfg = self.config
        batch, channels, frames, height, width = hidden_states.shape
        p_t, p_h, p_w = cfg.patch_size
        tokens = (frames // p_t) * (height // p_h) * (width // p_w)

        hs = hidden_states.to(dt)
        rows = ops.patchify(hs, (batch, channels, frames, height, width), hs.stride(), (p_t, p_h, p_w))
        text_state = self._text_state(encoder_hidden_states)
        for dedup in ("device", "unique"):
            conditioning = self._conditioning(timestep, batch, tokens, dedup)
            y = self.forward_rows(rows, batch, (frames, height, width), conditioning, text_state)
            if not self._dedup_overflowed():
                break  # (the second pass only runs for > 8 distinct per-token timesteps)

        c_out = y.shape[-1] // (p_t * p_h * p_w)
        out = torch.empty(batch, c_out, frames, height, width, dtype=dt, device=y.device)
        ops.unpatchify(y, out, (batch, c_out, frames, height, width), out.stride(), (p_t, p_h, p_w),
                       channel_last=True)  # :539-543
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)
