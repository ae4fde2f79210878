"""Native attention processors — the reference's ``AttnProcessor`` plugin surface (boundary #2).

Drop-in replacements for
  WanAttnProcessor2_0         reference architecture/transformer_wan.py:38-119
  CogVideoXAttnProcessor2_0   reference architecture/attention_processor.py:2805-2877 (and the Fused variant :2880-2948)
with the same call signatures, reading weights off the ``attn`` container (diffusers' ``Attention`` or
``frameino_b200.modules.Attention``). They issue only frameino_b200 CUDA kernels: fused-QKV tcgen05 GEMM, in-place
q/k norm + RoPE, tcgen05 flash attention, out-projection GEMM (optionally with the gated residual fused in).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def tensor_key(t: Optional[torch.Tensor]):
    """Identity of a tensor for the derived-weight caches: storage address, shape, dtype and the autograd in-place
    version counter. Inference tensors (created under ``torch.inference_mode()``) do not track a version — reading
    ``_version`` on them raises — so they are keyed by address alone.

    CONTRACT: the caches (concatenated q|k|v / k|v weights, stacked modulation tables, fp32 copies of norm affines,
    the per-prompt ``WanTextState``) see a change only if it bumps ``_version`` or moves the storage. Updates written
    through ``.data`` (``w.data.add_(delta)``, the usual LoRA-merge idiom) do neither: weights are FROZEN once
    ``model.prepare()`` or the first forward has run; after changing them in place call ``model.invalidate()``
    (or ``invalidate_derived(attn)`` for a single attention module on a foreign model)."""
    if t is None:
        return None
    return (t.data_ptr(), None if t.is_inference() else t._version, tuple(t.shape), t.dtype)


def invalidate_derived(module) -> None:
    """Drops the derived weights cached on ``module`` (an attention container the native processors have run on)."""
    module.__dict__.pop("_fino_cache", None)


def _fused_weights(attn, names: Tuple[str, ...], tag: str):
    """Concatenated [sum(out), in] weight (+bias) of several nn.Linear projections, cached on the module and
    refreshed when any source parameter changes (see ``tensor_key`` for what counts as a change)."""
    mods = [getattr(attn, n) for n in names]
    key = tuple((tensor_key(m.weight), tensor_key(m.bias)) for m in mods)
    cache = attn.__dict__.setdefault("_fino_cache", {})
    hit = cache.get(tag)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    with torch.no_grad():
        w = torch.cat([m.weight for m in mods], dim=0).contiguous()
        b = None
        if mods[0].bias is not None:
            b = torch.cat([m.bias for m in mods], dim=0).contiguous()
    cache[tag] = (key, w, b)
    return w, b


def _check_dtype(t: torch.Tensor, what: str) -> None:
    if t.dtype != torch.bfloat16:
        raise NotImplementedError(f"frameino_b200 kernels compute in bf16; {what} is {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(f"frameino_b200 has no CPU path; {what} is on {t.device}")


def _rope_table(t: torch.Tensor, head_dim: int) -> torch.Tensor:
    t = t.reshape(-1, head_dim)
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


class FinoWanAttnProcessor:
    """``processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None)`` -> Tensor.

    Extra keywords (only passed by frameino_b200's own block): ``fino_residual=(x, gate, row_index, rows_per_group)``
    fuses ``x + out * gate`` (transformer_wan.py:336 / :341) into the out-projection epilogue and returns ``x``;
    ``fino_text_kv=(k, v)`` supplies the cross-attention key (already normalised) and value projections of
    ``encoder_hidden_states`` computed once per prompt by ``project_text`` (they do not change over the sampler steps).
    """

    @staticmethod
    def project_text(attn, encoder_hidden_states: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(norm_k(to_k(text)), to_v(text))`` of a cross-attention module (transformer_wan.py:61-67), each
        [B, T, D] (views of one fused [B, T, 2D] GEMM output)."""
        _check_dtype(encoder_hidden_states, "encoder_hidden_states")
        w, bias = _fused_weights(attn, ("to_k", "to_v"), "kv")
        kv = ops.linear(encoder_hidden_states, w, bias)  # [B, T, 2D]
        d_model = w.shape[0] // 2
        k, v = kv[..., :d_model], kv[..., d_model:]
        if attn.norm_k is not None:
            ops.qk_norm_rope(k, attn.norm_k.weight, None, None, attn.heads, norm_mode=ops.QK_RMS_ACROSS_HEADS,
                             eps=getattr(attn.norm_k, "eps", 1e-6), rope_mode=ops.ROPE_NONE, seq_len=k.shape[1])
        return k, v

    def __call__(
        self,
        attn,
        hidden_states: torch.Tensor,
        encoder_hidden_states: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        fino_residual=None,
        fino_text_kv=None,
    ) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not supported (the reference never passes one)")
        if getattr(attn, "add_k_proj", None) is not None:
            raise NotImplementedError("Wan I2V added-KV branch (added_kv_proj_dim) is not part of the FrameINO path")
        _check_dtype(hidden_states, "hidden_states")
        heads = attn.heads
        b, n, _ = hidden_states.shape
        norm_q, norm_k = attn.norm_q, attn.norm_k
        eps = getattr(norm_q, "eps", 1e-6) if norm_q is not None else 1e-6
        if fino_text_kv is not None:  # cross-attention against pre-projected text K/V: only the query side is left
            if encoder_hidden_states is None or rotary_emb is not None:
                raise ValueError("fino_text_kv belongs to a cross-attention call (encoder_hidden_states, no rotary_emb)")
            k, v = fino_text_kv
            q = ops.linear(hidden_states, attn.to_q.weight, attn.to_q.bias)
            if norm_q is not None:
                ops.qk_norm_rope(q, norm_q.weight, None, None, heads, norm_mode=ops.QK_RMS_ACROSS_HEADS, eps=eps,
                                 rope_mode=ops.ROPE_NONE, seq_len=n)
            head_dim = q.shape[-1] // heads
            o = ops.attention(q, k, v, heads, scale=getattr(attn, "scale", head_dim ** -0.5))
            return self._out_proj(attn, o, fino_residual)
        if encoder_hidden_states is None:
            w, bias = _fused_weights(attn, ("to_q", "to_k", "to_v"), "qkv")
            qkv = ops.linear(hidden_states, w, bias)  # [B, N, 3D]
            d_model = w.shape[0] // 3
            q, k, v = qkv[..., :d_model], qkv[..., d_model:2 * d_model], qkv[..., 2 * d_model:]
        else:
            _check_dtype(encoder_hidden_states, "encoder_hidden_states")
            q = ops.linear(hidden_states, attn.to_q.weight, attn.to_q.bias)
            w, bias = _fused_weights(attn, ("to_k", "to_v"), "kv")
            kv = ops.linear(encoder_hidden_states, w, bias)  # [B, T, 2D]
            d_model = w.shape[0] // 2
            k, v = kv[..., :d_model], kv[..., d_model:]
        head_dim = d_model // heads
        scale = getattr(attn, "scale", head_dim ** -0.5)
        sp = attn.__dict__.get("_fino_sp") if encoder_hidden_states is None else None
        if sp is not None and sp.mode == "peer" and b == 1 and norm_q is not None and head_dim in (64, 128):
            # Ulysses over peer memory: norm + RoPE + 1st exchange in one kernel, attention + 2nd exchange in another
            cos = sin = None
            if rotary_emb is not None:
                cos, sin = _rope_table(rotary_emb[0], head_dim), _rope_table(rotary_emb[1], head_dim)
                if cos.shape[0] != n:
                    raise ValueError(f"rotary table has {cos.shape[0]} rows for {n} tokens")
            o = sp.fused_attention(qkv, norm_q.weight, norm_k.weight, heads, eps, cos, sin, scale)
            return self._out_proj(attn, o, fino_residual)
        if norm_q is not None or rotary_emb is not None:
            if norm_q is None:
                raise NotImplementedError("RoPE without qk-norm is not used by the reference Wan path")
            cos = sin = None
            mode = ops.ROPE_NONE
            if rotary_emb is not None:
                cos, sin = _rope_table(rotary_emb[0], head_dim), _rope_table(rotary_emb[1], head_dim)
                if cos.shape[0] != n:
                    raise ValueError(f"rotary table has {cos.shape[0]} rows for {n} tokens")
                mode = ops.ROPE_WAN
            ops.qk_norm_rope(q, norm_q.weight, k, norm_k.weight, heads, norm_mode=ops.QK_RMS_ACROSS_HEADS, eps=eps,
                             rope_mode=mode, cos=cos, sin=sin, seq_len=n)
        if sp is not None:
            o = sp.attention(qkv, heads, scale)  # Ulysses all-to-all (NCCL) around the full-sequence attention
        else:
            o = ops.attention(q, k, v, heads, scale=scale)
        return self._out_proj(attn, o, fino_residual)

    @staticmethod
    def _out_proj(attn, o: torch.Tensor, fino_residual):
        wo, bo = attn.to_out[0].weight, attn.to_out[0].bias
        if fino_residual is not None:
            x, gate, row_index, rows_per_group = fino_residual
            return ops.linear(o, wo, bo, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=gate, row_index=row_index,
                              rows_per_group=rows_per_group, out=x)
        return ops.linear(o, wo, bo)


class FinoCogVideoXAttnProcessor:
    """``processor(attn, hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None)``
    -> ``(hidden_states, encoder_hidden_states)`` (attention_processor.py:2815-2877).

    Extra keywords (frameino_b200's own block only): ``fino_joint_text_len=T`` means ``hidden_states`` already is the
    joint ``[text, video]`` sequence whose first T rows are text, and the joint output is returned un-split;
    ``fino_residual=(x, gate, row_index)`` fuses ``x + gate * out`` into the out-projection epilogue.
    """

    def __call__(
        self,
        attn,
        hidden_states: torch.Tensor,
        encoder_hidden_states: torch.Tensor,
        attention_mask: Optional[torch.Tensor] = None,
        image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        fino_joint_text_len: Optional[int] = None,
        fino_residual=None,
    ):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not supported (the reference never passes one)")
        _check_dtype(hidden_states, "hidden_states")
        if fino_joint_text_len is None:
            text_len = encoder_hidden_states.size(1)
            joint = torch.cat([encoder_hidden_states, hidden_states], dim=1)  # :2827
        else:
            text_len = fino_joint_text_len
            joint = hidden_states
        heads = attn.heads
        b, s, _ = joint.shape
        if getattr(attn, "fused_projections", False) and hasattr(attn, "to_qkv"):
            w, bias = attn.to_qkv.weight, attn.to_qkv.bias  # FusedCogVideoXAttnProcessor2_0, :2910
        else:
            w, bias = _fused_weights(attn, ("to_q", "to_k", "to_v"), "qkv")
        qkv = ops.linear(joint, w, bias)
        d_model = w.shape[0] // 3
        head_dim = d_model // heads
        q, k, v = qkv[..., :d_model], qkv[..., d_model:2 * d_model], qkv[..., 2 * d_model:]
        norm_q, norm_k = attn.norm_q, attn.norm_k
        cos = sin = None
        mode = ops.ROPE_NONE
        if image_rotary_emb is not None:
            cos, sin = _rope_table(image_rotary_emb[0], head_dim), _rope_table(image_rotary_emb[1], head_dim)
            if cos.shape[0] != s - text_len:
                raise ValueError(f"rotary table has {cos.shape[0]} rows for {s - text_len} video tokens")
            mode = ops.ROPE_COGVIDEOX
        scale = getattr(attn, "scale", head_dim ** -0.5)
        sp = attn.__dict__.get("_fino_sp")
        if sp is not None and fino_joint_text_len is None:
            raise NotImplementedError("sequence parallelism needs the native joint-sequence block")
        if (sp is not None and sp.mode == "peer" and norm_q is not None and head_dim == 64
                and not getattr(attn, "is_cross_attention", False)):
            # Ulysses over peer memory: per-head LayerNorm + RoPE + 1st exchange in one kernel per sample, attention +
            # 2nd exchange in another (the joint rows are sharded over the ranks; text rows sit on rank 0)
            o = sp.fused_attention_ln(qkv, norm_q, norm_k, heads, getattr(norm_q, "eps", 1e-6), cos, sin, text_len, scale)
        else:
            if norm_q is not None:
                ops.qk_norm_rope(q, norm_q.weight, k, norm_k.weight, heads, b0=norm_q.bias, b1=norm_k.bias,
                                 rope1=not getattr(attn, "is_cross_attention", False),
                                 norm_mode=ops.QK_LAYERNORM_PER_HEAD, eps=getattr(norm_q, "eps", 1e-6), rope_mode=mode,
                                 cos=cos, sin=sin, seq_len=s, rope_skip=text_len)
            elif mode != ops.ROPE_NONE:
                raise NotImplementedError("RoPE without qk-norm is not used by the reference CogVideoX path")
            if sp is not None:  # NCCL exchange: one all-to-all pair per batch element
                o = torch.cat([sp.attention(qkv[i:i + 1], heads, scale) for i in range(b)], dim=0)
            else:
                o = ops.attention(q, k, v, heads, scale=scale)
        if fino_residual is not None:  # x + gate * out (cogvideox_transformer_3d.py:146-147), joint layout only
            x, gate, row_index = fino_residual
            return ops.linear(o, attn.to_out[0].weight, attn.to_out[0].bias, epilogue=ops.EPI_GATE_RESIDUAL,
                              residual=x, gate=gate, row_index=row_index, round_product=True, out=x)
        out = ops.linear(o, attn.to_out[0].weight, attn.to_out[0].bias)  # :2870
        if fino_joint_text_len is not None:
            return out
        return out[:, text_len:], out[:, :text_len]  # :2874-2877 (video, text)
