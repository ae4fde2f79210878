"""Host-side mirror of the reference interface: state-dict layout, config surface, processor plumbing."""
import inspect
import os

import pytest
import torch

from frameino_b200 import synth
from frameino_b200.cogvideox import CogVideoXTransformer3DModel, sincos_pos_embed_3d
from frameino_b200.modules import Attention
from frameino_b200.processors import FinoCogVideoXAttnProcessor, FinoWanAttnProcessor
from frameino_b200.ulysses import SequenceParallel
from frameino_b200.wan import WanTransformer3DModel


def test_wan_state_dict_layout_is_the_diffusers_layout():
    m = WanTransformer3DModel(**synth.WAN_TINY)
    shapes = synth.wan_param_shapes(synth.WAN_TINY)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    # non-persistent rope buffers, as transformer_wan.py:225-226
    assert "rope.freqs_cos" not in sd and hasattr(m.rope, "freqs_cos")


def test_cog_state_dict_layout_is_the_diffusers_layout():
    m = CogVideoXTransformer3DModel(**synth.COG_TINY)
    shapes = synth.cog_param_shapes(synth.COG_TINY)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k


def test_forward_signatures_match_the_reference():
    wan = list(inspect.signature(WanTransformer3DModel.forward).parameters)
    assert wan == ["self", "hidden_states", "timestep", "encoder_hidden_states", "encoder_hidden_states_image",
                   "return_dict", "attention_kwargs"]  # transformer_wan.py:454-462
    cog = list(inspect.signature(CogVideoXTransformer3DModel.forward).parameters)
    assert cog == ["self", "hidden_states", "encoder_hidden_states", "timestep", "timestep_cond", "ofs",
                   "image_rotary_emb", "attention_kwargs", "return_dict"]  # cogvideox_transformer_3d.py:446-456
    wp = list(inspect.signature(FinoWanAttnProcessor.__call__).parameters)[:6]
    assert wp == ["self", "attn", "hidden_states", "encoder_hidden_states", "attention_mask", "rotary_emb"]
    cp = list(inspect.signature(FinoCogVideoXAttnProcessor.__call__).parameters)[:6]
    assert cp == ["self", "attn", "hidden_states", "encoder_hidden_states", "attention_mask", "image_rotary_emb"]


def test_config_and_pipeline_facing_attributes():
    m = WanTransformer3DModel(**synth.WAN_TINY)
    assert m.config.image_dim is None and m.config["patch_size"] == (1, 2, 2)
    assert m.dtype == torch.float32
    with m.cache_context("cond"):
        pass
    m.to_inference_dtype(torch.bfloat16)
    assert m.dtype == torch.bfloat16
    assert m.condition_embedder.time_embedder.linear_1.weight.dtype == torch.float32  # _keep_in_fp32_modules
    assert m.blocks[0].scale_shift_table.dtype == torch.float32
    assert m.blocks[0].norm2.weight.dtype == torch.float32
    assert m.blocks[0].attn1.norm_q.weight.dtype == torch.bfloat16  # norm_q is NOT kept in fp32
    c = CogVideoXTransformer3DModel(**synth.COG_TINY)
    for k in ("patch_size", "patch_size_t", "sample_width", "sample_height", "sample_frames", "attention_head_dim",
              "use_rotary_positional_embeddings", "ofs_embed_dim"):
        assert k in c.config


def test_processor_plumbing():
    m = CogVideoXTransformer3DModel(**synth.COG_TINY)
    procs = m.attn_processors
    assert len(procs) == synth.COG_TINY["num_layers"]
    assert all(k.endswith("attn1.processor") for k in procs)

    class Foreign:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, image_rotary_emb=None):
            return hidden_states, encoder_hidden_states

    m.set_attn_processor(Foreign())
    assert all(isinstance(p, Foreign) for p in m.attn_processors.values())
    with pytest.raises(ValueError):
        m.set_attn_processor({"x": Foreign()})
    m.fuse_qkv_projections()
    a = m.transformer_blocks[0].attn1
    assert a.fused_projections and a.to_qkv.weight.shape == (3 * 256, 256)
    assert torch.equal(a.to_qkv.weight[:256], a.to_q.weight)
    m.unfuse_qkv_projections()
    assert all(isinstance(p, Foreign) for p in m.attn_processors.values())


def test_attention_forward_filters_kwargs_by_processor_signature():
    seen = {}

    class P:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, rotary_emb=None):
            seen["rotary_emb"] = rotary_emb
            return hidden_states

    attn = Attention(64, 2, 32, "rms_norm_across_heads", 1e-6, processor=P())
    out = attn(torch.zeros(1, 4, 64), rotary_emb="R", not_a_param=1)  # extra kwarg is dropped with a warning
    assert out.shape == (1, 4, 64) and seen["rotary_emb"] == "R"
    with pytest.raises(ValueError):
        Attention(64, 2, 32, "l2", 1e-6)


def test_models_refuse_cpu_inputs():
    m = WanTransformer3DModel(**synth.WAN_TINY)
    h, ts, text = synth.make_wan_inputs(synth.WAN_TINY, 2, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(h, ts, text)


def test_unsupported_reference_branches_fail_loudly():
    with pytest.raises(NotImplementedError):
        WanTransformer3DModel(**{**synth.WAN_TINY, "image_dim": 1280})
    with pytest.raises(NotImplementedError):
        CogVideoXTransformer3DModel(**{**synth.COG_TINY, "patch_size_t": 2})


def test_sincos_pos_embed_matches_reference_golden(golden_dir):
    """The default (non-learned) pos_embedding buffer equals the reference's get_3d_sincos_pos_embed output: the
    reference golden run adds it to the patch embedding, so check through the value recipe instead — here only the
    closed form of one entry."""
    pe = sincos_pos_embed_3d(64, grid_w=4, grid_h=3, frames=2, spatial_scale=1.875, temporal_scale=1.0)
    assert pe.shape == (2, 12, 64)
    # temporal band = first 16 dims: [sin(t*w_i) (8), cos(t*w_i) (8)]
    t = 1.0
    w = 1.0 / 10000 ** (torch.arange(8, dtype=torch.float64) / 8.0)
    assert torch.allclose(pe[1, 0, :8].double(), torch.sin(t * w), atol=1e-6)
    assert torch.allclose(pe[1, 0, 8:16].double(), torch.cos(t * w), atol=1e-6)


def test_sequence_partition():
    assert SequenceParallel.partition(28160, 8) == (3520, 28160)
    assert SequenceParallel.partition(19126, 8) == (2391, 19128)
    assert SequenceParallel.partition(10, 4) == (3, 12)


def test_attention_wave_plan():
    """fino_attention_plan (host arithmetic of the attention launcher, include/frameino_b200.h): whole tiles first, the
    partly filled last wave split along KV so that it fits the idle SMs."""
    from frameino_b200 import ops

    # config 2 on 1 / 2 GPUs: 2640 / 1320 tiles on 148 SMs -> last wave 84 % / 92 % full, no split
    assert ops.attention_plan(28160, 28160, 24) == (2640, 1)
    assert ops.attention_plan(28160, 28160, 12) == (1320, 1)
    # 4-way / 8-way Ulysses: 660 = 4*148 + 68 -> 2 splits; 330 = 2*148 + 34 -> 4 splits
    assert ops.attention_plan(28160, 28160, 6) == (592, 2)
    assert ops.attention_plan(28160, 28160, 3) == (296, 4)
    # cross-attention (4 KV tiles): never split; split off; forced split clamps to the KV tile count
    assert ops.attention_plan(28160, 512, 24) == (2640, 1)
    assert ops.attention_plan(28160, 28160, 3, mode=0) == (330, 1)
    assert ops.attention_plan(1024, 384, 4, mode=7) == (0, 3)
    for nq, nk, h in [(39936, 39936, 3), (19126, 19126, 48), (19126, 19126, 6), (4096, 4096, 2)]:
        tiles = (nq + 255) // 256 * h
        n_full, s = ops.attention_plan(nq, nk, h)
        assert 0 <= n_full <= tiles and s >= 1
        if s > 1:
            assert n_full % 148 == 0 and (tiles - n_full) * s <= 148 and (nk + 127) // 128 // s >= 8
        else:
            assert n_full == tiles


def test_attention_wave_plan_per_head_dim():
    """fino_attention_plan_hd: head_dim 128 is the 256-row / 128-key decomposition, head_dim 64 the four-tile kernel's
    512-row tiles against 64-key steps (CogVideoX: 38 tiles x 48 heads = 1824 -> 12 full waves + 48 tiles split 3 ways)."""
    from frameino_b200 import _lib, ops

    assert ops.attention_plan_hd(28160, 28160, 3, 128) == ops.attention_plan(28160, 28160, 3) + (256,)
    n_full, s, rows = ops.attention_plan_hd(19126, 19126, 48, 64)
    assert rows == 512 and (n_full, s) == (1776, 3)
    assert ops.attention_plan_hd(19126, 19126, 48, 64, mode=0) == (1824, 1, 512)
    with pytest.raises(_lib.FinoError):
        ops.attention_plan_hd(100, 100, 1, 96)


def test_gemm_round_plan():
    """fino_gemm_plan: split-K only where the persistent pair kernel's last round is partly filled."""
    from frameino_b200 import ops

    # 8-way Ulysses shard, N = 3072: 14 x 12 = 168 tiles on 74 pairs -> 148 whole tiles + 20 tiles in 3 K-slices
    assert ops.gemm_plan(3520, 3072, 14336) == (148, 3)
    # ... but not when the K-slices would be too short to pay for the fix-up (measured: no gain at K = 3072)
    assert ops.gemm_plan(3520, 3072, 3072) == (168, 1)
    # last round more than half full: left alone (QKV / FFN-up at P = 8, everything at P = 1)
    assert ops.gemm_plan(3520, 9216, 3072) == (504, 1)
    assert ops.gemm_plan(3520, 14336, 3072) == (784, 1)
    assert ops.gemm_plan(28160, 3072, 3072) == (1320, 1)
    # K too short to slice; split off; forced mode on a small problem slices every tile
    assert ops.gemm_plan(3520, 3072, 512) == (168, 1)
    assert ops.gemm_plan(3520, 3072, 3072, mode=0) == (168, 1)
    assert ops.gemm_plan(512, 512, 2048, mode=4) == (0, 4)


def test_cache_context_keys_a_text_state_and_invalidates(monkeypatch):
    """model.cache_context(name) (pipeline_wan_i2v_motion_FrameINO.py:862, :873) keys the per-prompt text state: the
    bookkeeping (nesting, reuse, invalidation on in-place edits of the prompt or of a text-side weight, separate
    contexts) is host logic and runs on CPU with the device work stubbed out."""
    from frameino_b200.wan import WanTextState

    m = WanTransformer3DModel(**synth.WAN_TINY)
    calls = []

    def fake_prepare(ehs):
        calls.append(ehs)
        return WanTextState(text=ehs, kv=[None] * len(m.blocks), key=m._text_key(ehs), source=ehs)

    monkeypatch.setattr(m, "prepare_text", fake_prepare)
    a, b = torch.randn(1, 4, 64), torch.randn(1, 4, 64)
    assert m.__dict__.get("_fino_cache_name") is None
    s0 = m._text_state(a)
    assert m._text_state(a) is not s0 and len(calls) == 2          # outside a context nothing is kept
    with m.cache_context("cond"):
        s1 = m._text_state(a)
        assert m._text_state(a) is s1 and len(calls) == 3          # reused
        with m.cache_context("uncond"):
            assert m.__dict__["_fino_cache_name"] == "uncond"
            s2 = m._text_state(b)
            assert s2 is not s1 and m._text_state(b) is s2
        assert m.__dict__["_fino_cache_name"] == "cond"           # nesting restores the outer name
        assert m._text_state(a) is s1
        a.mul_(2.0)                                                # in-place edit of the prompt
        s3 = m._text_state(a)
        assert s3 is not s1 and m._text_state(a) is s3
        with torch.no_grad():
            m.blocks[1].attn2.to_k.weight.add_(1.0)               # text-side weight changed
        assert m._text_state(a) is not s3
        assert m._text_state(b) is not s2                          # same name, another prompt: recomputed
    assert m.__dict__.get("_fino_cache_name") is None
    with m.cache_context("uncond"):
        assert m._text_state(b) is not s2                          # the weight edit invalidated this context too
    m.clear_text_cache()
    assert m.__dict__["_fino_text_cache"] == {}


def test_sampler_helpers_on_cpu():
    from frameino_b200.sampling import flow_match_sigmas, wan_frameino_denoise_fused

    s = flow_match_sigmas(50, 5.0)
    assert s.shape == (51,) and float(s[0]) == 1.0 and float(s[-1]) == 0.0
    assert bool((s[:-1] > s[1:]).all())                            # strictly decreasing
    lin = torch.linspace(1.0, 1e-3, 50)
    assert torch.allclose(s[:-1], 5.0 * lin / (1.0 + 4.0 * lin))   # static shift 5 (train_wan_motion_FrameINO.yaml:43-50)
    m = WanTransformer3DModel(**synth.WAN_TINY)
    z = torch.zeros(1, 16, 2, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        wan_frameino_denoise_fused(m, z, z, torch.ones(1, 1, 2, 8, 8), torch.zeros(1, 16, 3, 8, 8),
                                   torch.zeros(1, 16, 1, 8, 8), torch.zeros(1, 4, 64), None, num_steps=1)


def test_moving_the_model_drops_derived_state():
    """.to() / accelerate offload hooks (reference app.py:163): fused weights, stacked tables, text states and gathered
    RoPE tables are derived from the parameters and must not outlive a move or a cast."""
    m = WanTransformer3DModel(**synth.WAN_TINY)
    attn = m.blocks[0].attn1
    attn.__dict__["_fino_cache"] = {"qkv": ("key", torch.zeros(1), None)}
    m._stacked_tables()
    m.rope(torch.zeros(1, 32, 2, 8, 8))
    m.__dict__["_fino_text_cache"] = {"cond": object()}
    assert m._sst_cache is not None and m.rope._cache
    m.to(torch.bfloat16)
    assert "_fino_cache" not in attn.__dict__ and m._sst_cache is None and m.rope._cache == {}
    assert m.__dict__["_fino_text_cache"] == {}
    c = CogVideoXTransformer3DModel(**synth.COG_TINY)
    blk = c.transformer_blocks[0]
    blk.norm1.affine_f32()
    assert blk.norm1._f32 is not None
    c.to(torch.bfloat16)
    assert blk.norm1._f32 is None


def test_joint_rope_sharding_for_cogvideox_sequence_parallel():
    """CogVideoX under Ulysses: joint rows [text | video] are cut into equal slices; each rank gets the RoPE rows of its
    own video tokens, rank 0 keeps the text rows un-rotated, the tail is zero-padded."""
    from frameino_b200.ulysses import shard_joint_rope

    text, n_video, hd = 3, 20, 4
    cos = torch.arange(float(n_video))[:, None].repeat(1, hd)
    sin = -cos
    for world in (1, 2, 4, 8):
        n_loc, n_pad = SequenceParallel.partition(text + n_video, world)
        seen = []
        for r in range(world):
            lt, c, s = shard_joint_rope(cos, sin, text, n_loc, r)
            assert lt == (text if r == 0 else 0) and c.shape == (n_loc - lt, hd) and torch.equal(s, -c)
            seen += c[:, 0].tolist()
        assert seen[:n_video] == list(range(n_video)) and all(v == 0 for v in seen[n_video:])
        assert len(seen) == n_pad - text
    with pytest.raises(NotImplementedError):
        shard_joint_rope(cos, sin, 9, 4, 0)
