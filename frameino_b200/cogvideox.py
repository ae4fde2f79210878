"""B200-native ``CogVideoXTransformer3DModel`` — drop-in for the reference class
(reference architecture/cogvideox_transformer_3d.py:164-562, FrameINO variant with ``use_FrameIn`` and
``extra_encoder_cond_channels``): same config, ``forward`` signature, diffusers state-dict keys and the attributes the
FrameINO CogVideoX pipeline reads (``config.*``, ``fuse_qkv_projections``).

Text and video tokens live in ONE joint ``[B, text+video, D]`` buffer for the whole forward (the reference re-cats
them in every block, cogvideox_transformer_3d.py:155, attention_processor.py:2827); the per-stream modulation of
``CogVideoXLayerNormZero`` becomes a per-token row index into a ``[2B, 3D]`` table.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from . import ops
from .modules import Attention, FeedForward, ModelBase, TimestepEmbedding, WeightOnlyNorm, compute_dtype, logger
from .processors import FinoCogVideoXAttnProcessor, tensor_key
from .wan import Transformer2DModelOutput


def _sincos_1d(dim: int, pos: torch.Tensor) -> torch.Tensor:
    omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
    out = torch.outer(pos.reshape(-1).double(), omega)
    return torch.cat([out.sin(), out.cos()], dim=1)


def sincos_pos_embed_3d(embed_dim: int, grid_w: int, grid_h: int, frames: int, spatial_scale: float,
                        temporal_scale: float) -> torch.Tensor:
    """Initial value of ``patch_embed.pos_embedding`` (embeddings.py:81-150): [frames, grid_h*grid_w, embed_dim]."""
    d_sp, d_t = 3 * embed_dim // 4, embed_dim // 4
    gh = torch.arange(grid_h, dtype=torch.float32) / spatial_scale
    gw = torch.arange(grid_w, dtype=torch.float32) / spatial_scale
    ww, hh = torch.meshgrid(gw, gh, indexing="xy")  # both [grid_h, grid_w]; "w goes first"
    sp = torch.cat([_sincos_1d(d_sp // 2, ww), _sincos_1d(d_sp // 2, hh)], dim=1)  # [gh*gw, d_sp]
    tt = _sincos_1d(d_t, torch.arange(frames, dtype=torch.float32) / temporal_scale)  # [frames, d_t]
    sp = sp[None].expand(frames, -1, -1)
    tt = tt[:, None, :].expand(-1, grid_h * grid_w, -1)
    return torch.cat([tt, sp], dim=-1).float()


class CogVideoXPatchEmbed(nn.Module):
    """Parameter layout of embeddings.py:626-715 (proj Conv2d, text_proj, pos_embedding buffer), 1.0 checkpoints."""

    def __init__(self, patch_size: int, in_channels: int, embed_dim: int, text_embed_dim: int, bias: bool,
                 sample_width: int, sample_height: int, sample_frames: int, temporal_compression_ratio: int,
                 max_text_seq_length: int, spatial_interpolation_scale: float, temporal_interpolation_scale: float,
                 use_positional_embeddings: bool, use_learned_positional_embeddings: bool, use_FrameIn: bool):
        super().__init__()
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        self.sample_height, self.sample_width, self.sample_frames = sample_height, sample_width, sample_frames
        self.temporal_compression_ratio = temporal_compression_ratio
        self.max_text_seq_length = max_text_seq_length
        self.use_positional_embeddings = use_positional_embeddings
        self.use_learned_positional_embeddings = use_learned_positional_embeddings
        self.use_FrameIn = use_FrameIn
        self.proj = nn.Conv2d(in_channels, embed_dim, kernel_size=(patch_size, patch_size), stride=patch_size, bias=bias)
        self.text_proj = nn.Linear(text_embed_dim, embed_dim)
        if use_positional_embeddings or use_learned_positional_embeddings:
            ph, pw = sample_height // patch_size, sample_width // patch_size
            frames = (sample_frames - 1) // temporal_compression_ratio + 1
            pos = sincos_pos_embed_3d(embed_dim, pw, ph, frames, spatial_interpolation_scale,
                                      temporal_interpolation_scale).flatten(0, 1)
            joint = pos.new_zeros(1, max_text_seq_length + pos.shape[0], embed_dim)
            joint[:, max_text_seq_length:] = pos
            self.register_buffer("pos_embedding", joint, persistent=use_learned_positional_embeddings)
        self._pos_cache: Dict[Any, torch.Tensor] = {}

    def positional_rows(self, text_len: int, frames: int, height: int, width: int, dtype) -> Optional[torch.Tensor]:
        """[text_len + frames*h/p*w/p, D] table added to the joint embedding (embeddings.py:754-803): FrameINO appends
        frame-0's rows for the ID frame (:772-775) and trilinearly resizes for non-default canvases (:782-798).
        Constant per canvas shape, so it is built once with torch ops and cached."""
        if not (self.use_positional_embeddings or self.use_learned_positional_embeddings):
            return None
        key = (text_len, frames, height, width, dtype, tensor_key(self.pos_embedding))
        hit = self._pos_cache.get(key)
        if hit is not None:
            return hit
        p = self.patch_size
        pos = self.pos_embedding
        tcr = self.temporal_compression_ratio
        pre_frames = (frames - 1) * tcr + 1
        post_frames = (self.sample_frames - 1) // tcr + 1
        ph, pw = self.sample_height // p, self.sample_width // p
        seq = height * width * frames // (p * p)
        if self.use_FrameIn:
            first = (pos.shape[1] - self.max_text_seq_length) // (frames - 1)
            pos = torch.cat([pos, pos[:, text_len:text_len + first]], dim=1)
        if self.sample_height != height or self.sample_width != width or self.sample_frames != pre_frames:
            if self.use_FrameIn:
                post_frames += 1
            d = pos.shape[-1]
            pv = pos[:, text_len:].reshape(1, post_frames, ph, pw, d).permute(0, 4, 1, 2, 3).float()
            pv = torch.nn.functional.interpolate(pv, size=[post_frames, height // p, width // p], mode="trilinear",
                                                 align_corners=False)
            pv = pv.permute(0, 2, 3, 4, 1).reshape(1, -1, d).to(pos.dtype)
            pos = torch.cat([pos[:, :text_len], pv], dim=1)[:, : text_len + seq]
        rows = pos[0].to(dtype).contiguous()
        if rows.shape[0] != text_len + seq:
            raise ValueError(f"positional table has {rows.shape[0]} rows for {text_len + seq} tokens")
        self._pos_cache = {key: rows}
        return rows


class LayerNormZero(nn.Module):
    """Parameter layout of diffusers CogVideoXLayerNormZero (upstream): linear (cond -> 6D), norm (LayerNorm D)."""

    def __init__(self, conditioning_dim: int, embedding_dim: int, elementwise_affine: bool, eps: float, chunks: int = 6):
        super().__init__()
        self.linear = nn.Linear(conditioning_dim, chunks * embedding_dim)
        self.norm = WeightOnlyNorm(embedding_dim, eps, elementwise_affine, bias=True)
        self._f32 = None

    def affine_f32(self):
        """fp32 copies of the LayerNorm gain/bias for the fused kernel (refreshed when the parameters change)."""
        w, b = self.norm.weight, self.norm.bias
        if w is None:
            return None, None
        key = (tensor_key(w), tensor_key(b))
        if self._f32 is None or self._f32[0] != key:
            self._f32 = (key, w.detach().float().contiguous(), b.detach().float().contiguous())
        return self._f32[1], self._f32[2]


class CogVideoXBlock(nn.Module):
    """cogvideox_transformer_3d.py:42-161 on the joint [B, text+video, D] buffer."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, time_embed_dim: int,
                 attention_bias: bool = False, qk_norm: bool = True, norm_elementwise_affine: bool = True,
                 norm_eps: float = 1e-5, ff_inner_dim: Optional[int] = None, ff_bias: bool = True,
                 attention_out_bias: bool = True):
        super().__init__()
        self.norm1 = LayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps)
        self.attn1 = Attention(dim, num_attention_heads, attention_head_dim, "layer_norm" if qk_norm else None, 1e-6,
                               bias=attention_bias, out_bias=attention_out_bias, processor=FinoCogVideoXAttnProcessor())
        self.norm2 = LayerNormZero(time_embed_dim, dim, norm_elementwise_affine, norm_eps)
        self.ff = FeedForward(dim, ff_inner_dim or 4 * dim, bias=ff_bias, final_dropout=True)
        self.eps = norm_eps

    def _mods(self, norm: LayerNormZero, emb_f32: torch.Tensor) -> torch.Tensor:
        """Linear(SiLU(temb)) -> fp32-of-bf16 [B, 6D], viewed as [2B, 3D]: row 2b = video (shift, scale, gate),
        row 2b+1 = text (enc_shift, enc_scale, enc_gate)."""
        m = ops.linear_small_m(emb_f32, norm.linear.weight, norm.linear.bias, act_in=1, round_in=True, round_out=True)
        return m.view(2 * m.shape[0], m.shape[1] // 2)

    def forward(self, joint: torch.Tensor, text_len: int, emb_f32: torch.Tensor, row_index: torch.Tensor,
                image_rotary_emb) -> torch.Tensor:
        dim = joint.shape[-1]
        tap = self.__dict__.pop("_fino_tap", None)  # parity tests: (dict, "transformer_blocks.i")
        # norm1 + attention + gated residual (:134-147)
        m1 = self._mods(self.norm1, emb_f32)
        g, b = self.norm1.affine_f32()
        h = ops.ln_modulate(joint, self.eps, gamma=g, beta=b, shift=m1[:, 0:dim], scale=m1[:, dim:2 * dim],
                            row_index=row_index, bf16_steps=True)
        if isinstance(self.attn1.processor, FinoCogVideoXAttnProcessor):
            joint = self.attn1(hidden_states=h, encoder_hidden_states=None, image_rotary_emb=image_rotary_emb,
                               fino_joint_text_len=text_len,
                               fino_residual=(joint, m1[:, 2 * dim:3 * dim], row_index))
        else:  # foreign processor: reference dataflow (split, call, re-join)
            a_v, a_t = self.attn1(hidden_states=h[:, text_len:], encoder_hidden_states=h[:, :text_len],
                                  image_rotary_emb=image_rotary_emb)
            a = torch.cat([a_t, a_v], dim=1).contiguous()
            joint = ops.gate_residual(joint, a, m1[:, 2 * dim:3 * dim], row_index=row_index, round_product=True,
                                      out=joint)
        if tap is not None:
            tap[0][tap[1] + ".after_attn"] = joint.clone()  # joint [text | video] rows after :146-147
        # norm2 + feed-forward + gated residual (:150-159)
        m2 = self._mods(self.norm2, emb_f32)
        g, b = self.norm2.affine_f32()
        h = ops.ln_modulate(joint, self.eps, gamma=g, beta=b, shift=m2[:, 0:dim], scale=m2[:, dim:2 * dim],
                            row_index=row_index, bf16_steps=True, out=h)
        up, down = self.ff.net[0].proj, self.ff.net[2]
        f = ops.linear(h, up.weight, up.bias, epilogue=ops.EPI_GELU_TANH)
        joint = ops.linear(f, down.weight, down.bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=joint,
                           gate=m2[:, 2 * dim:3 * dim], row_index=row_index, round_product=True, out=joint)
        return joint


class CogVideoXTransformer3DModel(ModelBase):
    """Drop-in for reference ``architecture.cogvideox_transformer_3d.CogVideoXTransformer3DModel``."""

    _supports_gradient_checkpointing = False
    _no_split_modules = ["CogVideoXBlock", "CogVideoXPatchEmbed"]

    def __init__(
        self,
        num_attention_heads: int = 30,
        attention_head_dim: int = 64,
        in_channels: int = 16,
        out_channels: Optional[int] = 16,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        time_embed_dim: int = 512,
        ofs_embed_dim: Optional[int] = None,
        text_embed_dim: int = 4096,
        num_layers: int = 30,
        dropout: float = 0.0,
        attention_bias: bool = True,
        sample_width: int = 90,
        sample_height: int = 60,
        sample_frames: int = 49,
        patch_size: int = 2,
        patch_size_t: Optional[int] = None,
        temporal_compression_ratio: int = 4,
        max_text_seq_length: int = 226,
        activation_fn: str = "gelu-approximate",
        timestep_activation_fn: str = "silu",
        norm_elementwise_affine: bool = True,
        norm_eps: float = 1e-5,
        spatial_interpolation_scale: float = 1.875,
        temporal_interpolation_scale: float = 1.0,
        use_rotary_positional_embeddings: bool = False,
        use_learned_positional_embeddings: bool = False,
        patch_bias: bool = True,
        extra_encoder_cond_channels: int = -1,
        use_FrameIn: bool = False,
    ):
        super().__init__()
        if patch_size_t is not None:
            raise NotImplementedError("patch_size_t (CogVideoX 1.5) is not part of the FrameINO CogVideoX-5B-I2V path")
        if ofs_embed_dim:
            raise NotImplementedError("ofs_embed_dim (CogVideoX 1.5 I2V) is not part of the FrameINO path")
        if activation_fn != "gelu-approximate" or timestep_activation_fn != "silu":
            raise NotImplementedError("only gelu-approximate / silu activations are used by the reference checkpoints")
        self._register_config(
            num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim, in_channels=in_channels,
            out_channels=out_channels, flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
            time_embed_dim=time_embed_dim, ofs_embed_dim=ofs_embed_dim, text_embed_dim=text_embed_dim,
            num_layers=num_layers, dropout=dropout, attention_bias=attention_bias, sample_width=sample_width,
            sample_height=sample_height, sample_frames=sample_frames, patch_size=patch_size, patch_size_t=patch_size_t,
            temporal_compression_ratio=temporal_compression_ratio, max_text_seq_length=max_text_seq_length,
            activation_fn=activation_fn, timestep_activation_fn=timestep_activation_fn,
            norm_elementwise_affine=norm_elementwise_affine, norm_eps=norm_eps,
            spatial_interpolation_scale=spatial_interpolation_scale,
            temporal_interpolation_scale=temporal_interpolation_scale,
            use_rotary_positional_embeddings=use_rotary_positional_embeddings,
            use_learned_positional_embeddings=use_learned_positional_embeddings, patch_bias=patch_bias,
            extra_encoder_cond_channels=extra_encoder_cond_channels, use_FrameIn=use_FrameIn,
        )
        inner_dim = num_attention_heads * attention_head_dim
        self.patch_embed = CogVideoXPatchEmbed(
            patch_size, in_channels, inner_dim, text_embed_dim, patch_bias, sample_width, sample_height, sample_frames,
            temporal_compression_ratio, max_text_seq_length, spatial_interpolation_scale, temporal_interpolation_scale,
            not use_rotary_positional_embeddings, use_learned_positional_embeddings, use_FrameIn)
        self.embedding_dropout = nn.Dropout(dropout)
        self.time_embedding = TimestepEmbedding(inner_dim, time_embed_dim)
        self.transformer_blocks = nn.ModuleList(
            [CogVideoXBlock(inner_dim, num_attention_heads, attention_head_dim, time_embed_dim,
                            attention_bias=attention_bias, norm_elementwise_affine=norm_elementwise_affine,
                            norm_eps=norm_eps) for _ in range(num_layers)]
        )
        self.norm_final = WeightOnlyNorm(inner_dim, norm_eps, norm_elementwise_affine, bias=True)
        self.norm_out = LayerNormZero(time_embed_dim, inner_dim, norm_elementwise_affine, norm_eps, chunks=2)
        self.proj_out = nn.Linear(inner_dim, patch_size * patch_size * out_channels)
        self.gradient_checkpointing = False
        self.original_attn_processors = None
        self._nf_f32 = None
        self.sequence_parallel = None  # set by frameino_b200.ulysses.enable_sequence_parallel (mode="nccl")

    def to_inference_dtype(self, dtype: torch.dtype = torch.bfloat16) -> "CogVideoXTransformer3DModel":
        return self.to(compute_dtype(dtype))  # float16 (reference app.py:156) -> bf16 with a warning

    def prepare(self) -> "CogVideoXTransformer3DModel":
        """Concatenated q|k|v projection weights, built once after loading instead of on the first forward."""
        from .processors import _fused_weights

        for b in self.transformer_blocks:
            if not (getattr(b.attn1, "fused_projections", False) and hasattr(b.attn1, "to_qkv")):
                _fused_weights(b.attn1, ("to_q", "to_k", "to_v"), "qkv")
        return self

    # cogvideox_transformer_3d.py:407-444 -------------------------------------------------------------------------
    def fuse_qkv_projections(self):
        self.original_attn_processors = self.attn_processors
        for module in self.modules():
            if isinstance(module, Attention):
                module.fuse_projections(fuse=True)
        self.set_attn_processor(FinoCogVideoXAttnProcessor())

    def unfuse_qkv_projections(self):
        if self.original_attn_processors is not None:
            self.set_attn_processor(self.original_attn_processors)
        for module in self.modules():
            if isinstance(module, Attention):
                module.fused_projections = False

    @torch.no_grad()
    def forward(
        self,
        hidden_states: torch.Tensor,
        encoder_hidden_states: torch.Tensor,
        timestep: Union[int, float, torch.Tensor],
        timestep_cond: Optional[torch.Tensor] = None,
        ofs: Optional[Union[int, float, torch.Tensor]] = None,
        image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        attention_kwargs: Optional[Dict[str, Any]] = None,
        return_dict: bool = True,
    ):
        if attention_kwargs is not None and attention_kwargs.get("scale", None) is not None:
            logger.warning("Passing `scale` via `attention_kwargs` when not using the PEFT backend is ineffective.")
        if timestep_cond is not None:
            raise NotImplementedError("timestep_cond is not used by the FrameINO pipelines")
        if not hidden_states.is_cuda:
            raise RuntimeError("frameino_b200 has no CPU path: move the model and inputs to a CUDA device")
        dt = self.proj_out.weight.dtype
        if dt != torch.bfloat16:
            raise NotImplementedError(f"model dtype {dt}: call .to_inference_dtype(torch.bfloat16) first "
                                      "(torch.float16 is accepted there and converted to bf16)")
        cfg = self.config
        batch, frames, channels, height, width = hidden_states.shape
        p = cfg.patch_size
        dim = cfg.num_attention_heads * cfg.attention_head_dim
        dev = hidden_states.device
        text_len = encoder_hidden_states.shape[1]
        n_video = frames * (height // p) * (width // p)
        seq = text_len + n_video

        # 1. time embedding (:478-485): sinusoid fp32 -> model dtype -> Linear, SiLU, Linear
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=dev)
        ts = timestep.to(dev).reshape(-1).float()
        if ts.numel() == 1 and batch > 1:
            ts = ts.expand(batch)
        ts = ts.contiguous()
        if ts.numel() != batch or batch > 8:
            raise NotImplementedError("timestep must hold one value per sample (batch <= 8)")
        t_emb = ops.timestep_embedding(ts, dim, cfg.flip_sin_to_cos, float(cfg.freq_shift))
        te = self.time_embedding
        e1 = ops.linear_small_m(t_emb, te.linear_1.weight, te.linear_1.bias, act_out=1, round_in=True, round_out=True)
        emb = ops.linear_small_m(e1, te.linear_2.weight, te.linear_2.bias, round_out=True)  # fp32-of-bf16 [B, te]

        # 2. patch embedding into the joint buffer, positional rows fused as the GEMM residual (:494)
        pe = self.patch_embed
        joint = torch.empty(batch, seq, dim, dtype=dt, device=dev)
        pos = pe.positional_rows(text_len, frames, height, width, dt)
        hs = hidden_states.to(dt)
        st = hs.stride()
        rows = ops.patchify(hs, (batch, channels, frames, height, width), (st[0], st[2], st[1], st[3], st[4]), (1, p, p))
        text_in = encoder_hidden_states.to(dt)
        wv = pe.proj.weight.view(dim, -1)
        for b in range(batch):
            kw_t = dict(epilogue=ops.EPI_GATE_RESIDUAL, residual=pos[:text_len]) if pos is not None else {}
            kw_v = dict(epilogue=ops.EPI_GATE_RESIDUAL, residual=pos[text_len:]) if pos is not None else {}
            ops.linear(text_in[b], pe.text_proj.weight, pe.text_proj.bias, out=joint[b, :text_len], **kw_t)
            ops.linear(rows[b * n_video:(b + 1) * n_video], wv, pe.proj.bias, out=joint[b, text_len:], **kw_v)

        # per-token modulation row: 2b for video tokens, 2b+1 for text tokens
        ar = torch.arange(seq, device=dev)
        row_index = (2 * torch.arange(batch, device=dev)[:, None] + (ar[None, :] < text_len)).to(torch.int32)
        row_index = row_index.reshape(-1).contiguous()

        # Ulysses (frameino_b200/ulysses.py): every rank keeps a contiguous slice of the JOINT rows; everything except
        # the attention core is row-local. Text rows sit at the front, i.e. on rank 0.
        sp = self.sequence_parallel
        seq_loc, text_loc, rope_loc = seq, text_len, image_rotary_emb
        if sp is not None:
            from .ulysses import shard_joint_rope

            sp.plan(seq)
            seq_loc = sp.n_loc
            joint = sp.shard_rows(joint).contiguous()
            row_index = sp.shard_rows(row_index.view(batch, seq)).reshape(-1).contiguous()
            if image_rotary_emb is not None:
                text_loc, c_loc, s_loc = shard_joint_rope(image_rotary_emb[0].reshape(-1, cfg.attention_head_dim).float(),
                                                          image_rotary_emb[1].reshape(-1, cfg.attention_head_dim).float(),
                                                          text_len, seq_loc, sp.rank)
                rope_loc = (c_loc, s_loc)
            else:
                text_loc = text_len if sp.rank == 0 else 0

        # 3. transformer blocks (:503-529)
        taps = self.__dict__.get("_fino_taps")  # parity tests set this to a dict to collect per-layer outputs
        if taps is not None:
            taps["patch_embed"] = joint.clone()
        for i, block in enumerate(self.transformer_blocks):
            if taps is not None:
                block.__dict__["_fino_tap"] = (taps, f"transformer_blocks.{i}")
            joint = block(joint, text_loc, emb, row_index, rope_loc)
            if taps is not None:
                taps[f"transformer_blocks.{i}.out"] = joint[:, text_loc:].clone()
                taps[f"transformer_blocks.{i}.enc"] = joint[:, :text_loc].clone()

        # 4. final norms + projection (:531-542)
        nf = self.norm_final
        if nf.weight is not None:
            key = (tensor_key(nf.weight), tensor_key(nf.bias))
            if self._nf_f32 is None or self._nf_f32[0] != key:
                self._nf_f32 = (key, nf.weight.detach().float().contiguous(), nf.bias.detach().float().contiguous())
            g, bta = self._nf_f32[1], self._nf_f32[2]
        else:
            g = bta = None
        h = ops.ln_modulate(joint, cfg.norm_eps, gamma=g, beta=bta)
        no = self.norm_out
        m = ops.linear_small_m(emb, no.linear.weight, no.linear.bias, act_in=1, round_in=True, round_out=True)  # [B, 2D]
        g2, b2 = no.affine_f32()
        h = ops.ln_modulate(h, cfg.norm_eps, gamma=g2, beta=b2, shift=m[:, :dim], scale=m[:, dim:],
                            rows_per_group=seq_loc, bf16_steps=True, out=h)
        c_out = self.proj_out.weight.shape[0] // (p * p)
        out = torch.empty(batch, frames, c_out, height, width, dtype=dt, device=dev)
        for b in range(batch):
            if sp is not None:  # project the local rows, gather the joint sequence, drop the text rows
                y_loc = ops.linear(h[b], self.proj_out.weight, self.proj_out.bias)
                y = sp.gather_rows(y_loc[None])[0, text_len:].contiguous()
            else:
                y = ops.linear(h[b, text_len:], self.proj_out.weight, self.proj_out.bias)
            ob = out[b:b + 1]
            so = ob.stride()
            ops.unpatchify(y, ob, (1, c_out, frames, height, width), (so[0], so[2], so[1], so[3], so[4]), (1, p, p),
                           channel_last=False)  # :549-550
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)
