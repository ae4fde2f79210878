"""Parameter containers that mirror the diffusers modules the reference builds its transformers from, so that a
diffusers-format state dict loads unchanged and the ``AttnProcessor`` plugin surface keeps working without diffusers
being installed. They hold parameters and dispatch; arithmetic lives in the CUDA kernels (``frameino_b200.ops``).

Mirrors (interface only, re-implemented):
  Attention            reference architecture/attention_processor.py:50-297 (container), :522-600 (processor plumbing)
  RMSNorm/FP32LayerNorm  diffusers.models.normalization (upstream; parameters only)
  FeedForward          diffusers.models.attention.FeedForward (upstream; keys net.0.proj / net.2)
"""
from __future__ import annotations

import inspect
import logging
from contextlib import contextmanager
from typing import Any, Dict, Optional

import torch
from torch import nn

logger = logging.getLogger("frameino_b200")


class ConfigDict(dict):
    """Attribute-and-item access config, like diffusers' FrozenDict (the pipelines read ``model.config.xxx``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is read-only")


class WeightOnlyNorm(nn.Module):
    """Holds ``weight`` (and optionally ``bias``) of a norm layer; the math is fused into kernels."""

    def __init__(self, dim: int, eps: float, elementwise_affine: bool = True, bias: bool = False):
        super().__init__()
        self.dim = dim
        self.eps = eps
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(dim))
            self.bias = nn.Parameter(torch.zeros(dim)) if bias else None
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, *a, **k):  # pragma: no cover - never on the native path
        raise RuntimeError("WeightOnlyNorm is a parameter holder; its arithmetic is fused into frameino_b200 kernels")


class GELUProj(nn.Module):
    """``net.0`` of diffusers FeedForward: holds ``proj``."""

    def __init__(self, dim_in: int, dim_out: int, bias: bool = True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)


class FeedForward(nn.Module):
    """Parameter layout of diffusers FeedForward(activation_fn="gelu-approximate"): net.0.proj, net.1 dropout, net.2."""

    def __init__(self, dim: int, inner_dim: int, dim_out: Optional[int] = None, bias: bool = True,
                 final_dropout: bool = False):
        super().__init__()
        dim_out = dim_out or dim
        mods = [GELUProj(dim, inner_dim, bias), nn.Dropout(0.0), nn.Linear(inner_dim, dim_out, bias=bias)]
        if final_dropout:
            mods.append(nn.Dropout(0.0))
        self.net = nn.ModuleList(mods)


class Attention(nn.Module):
    """Container with the attribute surface a diffusers ``AttnProcessor`` reads (attention_processor.py:50-297):
    ``to_q/to_k/to_v/to_out/norm_q/norm_k/heads/scale/add_k_proj/...`` plus ``set_processor/get_processor`` and a
    ``forward`` that filters kwargs by the processor's signature (attention_processor.py:583-600)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, qk_norm: Optional[str], eps: float, bias: bool = True,
                 out_bias: bool = True, processor: Any = None, cross_attention_dim: Optional[int] = None):
        super().__init__()
        self.inner_dim = heads * dim_head
        self.query_dim = query_dim
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.use_bias = bias
        self.is_cross_attention = cross_attention_dim is not None
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.fused_projections = False
        self.qk_norm = qk_norm
        self.eps = eps
        if qk_norm is None:
            self.norm_q = None
            self.norm_k = None
        elif qk_norm == "rms_norm_across_heads":  # attention_processor.py:208-211
            self.norm_q = WeightOnlyNorm(self.inner_dim, eps)
            self.norm_k = WeightOnlyNorm(self.inner_dim, eps)
        elif qk_norm == "layer_norm":  # attention_processor.py:195-197
            self.norm_q = WeightOnlyNorm(dim_head, eps, bias=True)
            self.norm_k = WeightOnlyNorm(dim_head, eps, bias=True)
        else:
            raise ValueError(f"unknown qk_norm: {qk_norm}. Supported here: None, 'layer_norm', 'rms_norm_across_heads'")
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        # branches that exist on the diffusers class and are dead for Wan2.2 / CogVideoX-5B (added_kv_proj_dim=None)
        self.add_q_proj = None
        self.add_k_proj = None
        self.add_v_proj = None
        self.norm_added_q = None
        self.norm_added_k = None
        self.to_add_out = None
        self.processor = None
        self.set_processor(processor)

    def set_processor(self, processor: Any) -> None:
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False) -> Any:
        return self.processor

    @torch.no_grad()
    def fuse_projections(self, fuse: bool = True) -> None:
        """attention_processor.py:770-820: concatenate q/k/v (self) or k/v (cross) weights into to_qkv / to_kv."""
        device, dtype = self.to_q.weight.device, self.to_q.weight.dtype
        if not self.is_cross_attention:
            w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight])
            self.to_qkv = nn.Linear(w.shape[1], w.shape[0], bias=self.use_bias, device=device, dtype=dtype)
            self.to_qkv.weight.copy_(w)
            if self.use_bias:
                self.to_qkv.bias.copy_(torch.cat([self.to_q.bias, self.to_k.bias, self.to_v.bias]))
        else:
            w = torch.cat([self.to_k.weight, self.to_v.weight])
            self.to_kv = nn.Linear(w.shape[1], w.shape[0], bias=self.use_bias, device=device, dtype=dtype)
            self.to_kv.weight.copy_(w)
            if self.use_bias:
                self.to_kv.bias.copy_(torch.cat([self.to_k.bias, self.to_v.bias]))
        self.fused_projections = fuse

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        params = set(inspect.signature(self.processor.__call__).parameters.keys())
        unused = [k for k in cross_attention_kwargs if k not in params and k not in ("ip_adapter_masks", "ip_hidden_states")]
        if unused:
            logger.warning("cross_attention_kwargs %s are not expected by %s and will be ignored.", unused,
                           self.processor.__class__.__name__)
        kwargs = {k: v for k, v in cross_attention_kwargs.items() if k in params}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class TimestepEmbedding(nn.Module):
    """Parameter layout of embeddings.py:1317-1362 (linear_1, SiLU, linear_2)."""

    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class TextProjection(nn.Module):
    """Parameter layout of PixArtAlphaTextProjection, embeddings.py:2247-2273 (linear_1, GELU-tanh, linear_2)."""

    def __init__(self, in_features: int, hidden_size: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.linear_2 = nn.Linear(hidden_size, hidden_size)


def compute_dtype(requested: torch.dtype) -> torch.dtype:
    """The dtype the kernels will run a model in when ``requested`` is asked for. The sm_100a kernels compute in bf16
    (fp32 accumulation and fp32 islands). The reference demo loads its transformer with ``torch_dtype=torch.float16``
    (app.py:156): that request is honoured by CONVERTING to bf16, loudly — same 16-bit storage, 8 instead of 11
    significand bits, fp32's exponent range (no fp16 overflow handling needed); ``model.dtype`` then reports bfloat16,
    which is what the pipeline casts its inputs to (pipeline_wan_i2v_motion_FrameINO.py:746). fp32 is refused."""
    if requested == torch.bfloat16:
        return requested
    if requested == torch.float16:
        logger.warning("frameino_b200 computes in bfloat16: torch_dtype=float16 weights are converted to bfloat16 "
                       "(model.dtype reports bfloat16; outputs are bfloat16)")
        return torch.bfloat16
    raise NotImplementedError(f"frameino_b200 kernels compute in bf16; cannot run the model in {requested} "
                              "(use torch.bfloat16, or torch.float16 which is converted to bf16 with a warning)")


class ModelBase(nn.Module):
    """The slice of diffusers ModelMixin/ConfigMixin/CacheMixin that the FrameINO pipelines touch."""

    config: ConfigDict

    def _register_config(self, **kwargs) -> None:
        object.__setattr__(self, "config", ConfigDict(kwargs))

    @property
    def dtype(self) -> torch.dtype:
        # the dtype of the GEMM weights (fp32 "keep" modules are skipped, as diffusers' get_parameter_dtype does)
        for n, p in self.named_parameters():
            if p.is_floating_point() and not any(k in n for k in getattr(self, "_keep_in_fp32_modules", []) or []):
                return p.dtype
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @contextmanager
    def cache_context(self, name: str):
        """diffusers CacheMixin.cache_context: a no-op unless a cache hook is enabled (none here)."""
        yield

    def _apply(self, fn, *args, **kwargs):
        """``.to()`` / ``.cuda()`` / ``.cpu()`` (and accelerate's offload hooks, reference app.py:163) move or cast the
        parameters; everything derived from them (fused projection weights, stacked modulation tables, per-prompt text
        state, gathered RoPE tables) is dropped so that it is rebuilt from the new tensors and the old device memory is
        released."""
        out = super()._apply(fn, *args, **kwargs)
        self.drop_derived_state()
        return out

    def invalidate(self) -> "ModelBase":
        """Call after changing weights IN PLACE through ``.data`` (e.g. merging a LoRA with ``w.data.add_(delta)``):
        such writes neither move the storage nor bump the version counter the derived-weight caches are keyed on
        (``processors.tensor_key``), so the concatenated projections / stacked tables / per-prompt text state would be
        stale. Weights are otherwise treated as frozen once ``prepare()`` or the first forward has run."""
        self.drop_derived_state()
        return self

    def drop_derived_state(self) -> None:
        for mod in self.modules():
            d = mod.__dict__
            d.pop("_fino_cache", None)
            if "_sst_cache" in d:
                d["_sst_cache"] = None
            if "_fino_text_cache" in d:
                d["_fino_text_cache"] = {}
            if isinstance(d.get("_cache"), dict):
                d["_cache"] = {}
            if isinstance(d.get("_pos_cache"), dict):
                d["_pos_cache"] = {}
            for k in ("_nf_f32", "_f32"):
                if k in d:
                    d[k] = None

    # ---- checkpoints (frameino_b200/loading.py; reference app.py:150-156 builds its transformer the same way) -------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None,
                        torch_dtype: torch.dtype = torch.bfloat16, device=None, **kwargs):
        """Loads a diffusers-format model directory (``config.json`` + ``diffusion_pytorch_model*.safetensors``, single
        or sharded) with the ``torch_dtype`` + ``_keep_in_fp32_modules`` dtype policy, directly onto ``device``."""
        from . import loading

        return loading.from_pretrained(cls, pretrained_model_name_or_path, subfolder=subfolder, torch_dtype=torch_dtype,
                                       device=device, **kwargs)

    def save_pretrained(self, save_directory: str, max_shard_size: int = 10 << 30):
        from . import loading

        return loading.save_pretrained(self, save_directory, max_shard_bytes=max_shard_size)

    def prepare(self) -> "ModelBase":
        """Builds the derived weights the kernels read (concatenated projection matrices, cached per attention module)
        ahead of the first forward. Overridden per model; harmless to skip (they are built lazily otherwise)."""
        return self

    def enable_gradient_checkpointing(self):  # inference-only build
        raise NotImplementedError("frameino_b200 is an inference path; training/backward is out of scope")

    @property
    def attn_processors(self) -> Dict[str, Any]:
        procs = {}
        for name, mod in self.named_modules():
            if hasattr(mod, "get_processor"):
                procs[f"{name}.processor"] = mod.get_processor()
        return procs

    def set_attn_processor(self, processor) -> None:
        """cogvideox_transformer_3d.py:372-404: one processor for all layers or a dict keyed '<module>.processor'."""
        count = len(self.attn_processors)
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(
                f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                f" number of attention layers: {count}. Please make sure to pass {count} processor classes."
            )
        for name, mod in self.named_modules():
            if hasattr(mod, "set_processor"):
                mod.set_processor(processor.pop(f"{name}.processor") if isinstance(processor, dict) else processor)
