"""Checkpoint loading: diffusers-format directories (config.json + single / sharded safetensors)
into the native models, on CPU (construction and loading need no GPU; only forward does)."""
import json
import os

import pytest
import torch

from frameino_b200 import loading, synth
from frameino_b200.cogvideox import CogVideoXTransformer3DModel
from frameino_b200.wan import WanTransformer3DModel


def _write_ckpt(tmp_path, cfg, sd, class_name, shard_bytes=None):
    from safetensors.torch import save_file

    d = tmp_path / "transformer"
    d.mkdir()
    meta = {"_class_name": class_name, "_diffusers_version": "0.35.0.dev0"}
    meta.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()})
    (d / "config.json").write_text(json.dumps(meta))
    if shard_bytes is None:
        save_file({k: v.contiguous() for k, v in sd.items()}, str(d / loading.WEIGHTS_NAME))
    else:
        shards = loading.shard_state_dict({k: v.contiguous() for k, v in sd.items()}, shard_bytes)
        assert len(shards) > 1
        wm = {}
        for i, sh in enumerate(shards):
            fn = f"diffusion_pytorch_model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
            save_file(sh, str(d / fn))
            wm.update({k: fn for k in sh})
        (d / loading.INDEX_NAME).write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    return str(d)


@pytest.mark.parametrize("sharded", [False, True])
def test_wan_from_pretrained_dtype_policy_and_values(tmp_path, sharded):
    cfg = synth.WAN_TINY
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.float32)  # an fp32 checkpoint, like the released ones
    sd["blocks.0.attn2.norm_added_q.weight"] = torch.ones(4)          # listed in _keys_to_ignore_on_load_unexpected
    path = _write_ckpt(tmp_path, cfg, sd, "WanTransformer3DModel", shard_bytes=200_000 if sharded else None)
    model = WanTransformer3DModel.from_pretrained(path, torch_dtype=torch.bfloat16, device="cpu")
    assert not model.training and model.dtype == torch.bfloat16
    assert model.config.num_layers == cfg["num_layers"] and tuple(model.config.patch_size) == tuple(cfg["patch_size"])
    got = model.state_dict()
    for k, v in sd.items():
        if "norm_added_q" in k:
            assert k not in got
            continue
        keep = any(m in k for m in WanTransformer3DModel._keep_in_fp32_modules)
        assert got[k].dtype == (torch.float32 if keep else torch.bfloat16), k
        assert torch.equal(got[k], v.to(got[k].dtype)), k
    assert model.rope.freqs_cos.dtype == torch.float32 and model.rope.freqs_cos.device.type == "cpu"
    # same result as the explicit path the GPU tests use
    ref = WanTransformer3DModel(**cfg)
    ref.load_state_dict({k: v for k, v in sd.items() if "norm_added_q" not in k})
    ref = ref.to_inference_dtype(torch.bfloat16)
    for (k, a), (_, b) in zip(sorted(got.items()), sorted(ref.state_dict().items())):
        assert a.dtype == b.dtype and torch.equal(a, b), k


def test_cog_from_pretrained_and_subfolder(tmp_path):
    cfg = synth.COG_TINY
    sd = synth.make_cog_state_dict(cfg, seed=0, dtype=torch.float32)
    _write_ckpt(tmp_path, cfg, sd, "CogVideoXTransformer3DModel")
    model = CogVideoXTransformer3DModel.from_pretrained(str(tmp_path), subfolder="transformer", device="cpu")
    got = model.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert got[k].dtype == torch.bfloat16 and torch.equal(got[k], v.bfloat16()), k


def test_save_pretrained_round_trip_sharded(tmp_path):
    cfg = synth.WAN_TINY
    model = WanTransformer3DModel(**cfg)
    model.load_state_dict(synth.make_wan_state_dict(cfg, seed=1, dtype=torch.float32))
    model = model.to_inference_dtype(torch.bfloat16)
    files = model.save_pretrained(str(tmp_path / "out"), max_shard_size=150_000)
    assert len(files) > 1 and os.path.exists(tmp_path / "out" / loading.INDEX_NAME)
    again = WanTransformer3DModel.from_pretrained(str(tmp_path / "out"), device="cpu")
    for (k, a), (_, b) in zip(sorted(model.state_dict().items()), sorted(again.state_dict().items())):
        assert a.dtype == b.dtype and torch.equal(a, b), k


def test_mismatched_checkpoints_fail_loudly(tmp_path):
    cfg = synth.WAN_TINY
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.float32)
    bad = dict(sd)
    bad.pop("proj_out.bias")
    bad["blocks.0.made_up.weight"] = torch.zeros(3)
    path = _write_ckpt(tmp_path, cfg, bad, "WanTransformer3DModel")
    with pytest.raises(ValueError, match="missing keys.*proj_out.bias.*unexpected keys.*made_up"):
        WanTransformer3DModel.from_pretrained(path, device="cpu")
    with pytest.raises(FileNotFoundError):
        WanTransformer3DModel.from_pretrained(str(tmp_path / "nope"), device="cpu")
    (tmp_path / "b").mkdir()
    shaped = dict(sd)
    shaped["proj_out.bias"] = torch.zeros(sd["proj_out.bias"].numel() + 1)
    path2 = _write_ckpt(tmp_path / "b", cfg, shaped, "WanTransformer3DModel")
    with pytest.raises(ValueError, match="proj_out.bias: checkpoint shape"):
        WanTransformer3DModel.from_pretrained(path2, device="cpu")


def test_no_init_weights_restores_constructors():
    import torch.nn as nn

    before = nn.Linear.reset_parameters
    with loading.no_init_weights():
        assert nn.Linear.reset_parameters is not before
    assert nn.Linear.reset_parameters is before


@pytest.mark.gpu
def test_from_pretrained_on_gpu_matches_explicit_load(tmp_path):
    """Loaded straight onto the GPU, prepared, and the forward is bit-identical to the state-dict path of the other
    GPU tests."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    cfg = synth.WAN_SMALL
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.float32)
    path = _write_ckpt(tmp_path, cfg, sd, "WanTransformer3DModel", shard_bytes=4_000_000)
    model = WanTransformer3DModel.from_pretrained(path)  # defaults: bf16, current CUDA device
    assert all(p.is_cuda for p in model.parameters()) and model.rope.freqs_cos.is_cuda
    assert "qkv" in model.blocks[0].attn1.__dict__["_fino_cache"] and "kv" in model.blocks[0].attn2.__dict__["_fino_cache"]
    ref = WanTransformer3DModel(**cfg)
    ref.load_state_dict(sd)
    ref = ref.to_inference_dtype(torch.bfloat16).cuda().eval()
    hidden, ts, text = synth.make_wan_inputs(cfg, 3, 16, 16, n_id=1, text_len=16, text_true_len=11, dtype=torch.bfloat16)
    args = dict(hidden_states=hidden.cuda(), timestep=ts.cuda(), encoder_hidden_states=text.cuda(), return_dict=False)
    assert torch.equal(model(**args)[0], ref(**args)[0])


def test_vae_from_pretrained_keeps_fp32_and_packs_bf16(tmp_path):
    """The reference loads its VAE with torch_dtype=float32 (app.py:157): parameters stay fp32 under the diffusers key
    names, the packed bf16 conv weights the kernels read are derived (tap-major, channels padded to 64 per tap)."""
    from safetensors.torch import save_file

    from frameino_b200.vae import AutoencoderKLWan, ConvParams

    cfg = dict(synth.VAE_TINY)
    sd = synth.make_vae_state_dict(cfg, seed=0)
    d = tmp_path / "vae"
    d.mkdir()
    meta = {"_class_name": "AutoencoderKLWan", "latents_mean": [0.0] * 16, "latents_std": [1.0] * 16}
    meta.update(cfg)
    (d / "config.json").write_text(json.dumps(meta))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / loading.WEIGHTS_NAME))
    vae = AutoencoderKLWan.from_pretrained(str(tmp_path), subfolder="vae", torch_dtype=torch.float32, device="cpu")
    assert vae.dtype == torch.float32 and vae.config.z_dim == 16 and vae.config.latents_std == [1.0] * 16
    got = vae.state_dict()
    assert set(got) == set(sd)
    for k, v in sd.items():
        assert got[k].dtype == torch.float32 and torch.equal(got[k], v), k
    conv = vae.decoder.up_blocks[2].resnets[0].conv1  # 256 -> 128 at the tiny widths
    wp, bp = conv.packed()
    assert isinstance(conv, ConvParams) and wp.dtype == torch.bfloat16 and wp.shape == (128, 27 * 256)
    w = conv.weight.detach()
    assert torch.equal(wp.view(128, 27, 256)[:, 13, :], w[:, :, 1, 1, 1].to(torch.bfloat16))  # centre tap
    c_in = vae.decoder.conv_in  # z 16 -> 256: channels padded 16 -> 64 per tap with zeros
    wp, _ = c_in.packed()
    assert wp.shape == (256, 27 * 64) and float(wp.view(256, 27, 64)[:, :, 16:].abs().max()) == 0.0
    c_out = vae.decoder.conv_out  # 64 -> 12: output rows padded 12 -> 16
    wp, bp = c_out.packed()
    assert wp.shape[0] == 16 and float(wp[12:].abs().max()) == 0.0 and float(bp[12:].abs().max()) == 0.0
    with pytest.raises(NotImplementedError):
        AutoencoderKLWan(**{**cfg, "is_residual": False})
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(1, 16, 1, 4, 6))  # no CPU path
