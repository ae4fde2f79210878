"""CPU: the pipeline restatement (oracle/pipeline_oracle.py) against the pieces it is made of, and the host-side
contract of frameino_b200.pipeline (validation errors, pre / post-processing, no CPU path)."""
import numpy as np
import pytest
import torch

from frameino_b200 import synth
from frameino_b200.pipeline import VideoProcessor, WanFrameINOPipeline, retrieve_latents


def _tiny():
    from oracle import wan_oracle

    vcfg = synth.with_latent_stats(synth.VAE_TINY)
    vsd = synth.make_vae_state_dict(synth.VAE_TINY, seed=1)
    wcfg = synth.WAN_TINY
    wsd = synth.make_wan_state_dict(wcfg, seed=0)
    return vcfg, vsd, wan_oracle.WanConfig(**wcfg), wsd


def test_oracle_prepare_latents_shapes_and_stats():
    """pipeline :400-536: shapes of the five returned tensors, the reciprocal-std normalisation, the zero trajectory
    padding on the ID frames, the first-frame mask."""
    from oracle import pipeline_oracle, vae_oracle

    vcfg, vsd, _, _ = _tiny()
    inp = synth.make_pipeline_inputs(vcfg, 64, num_frames=9, height=64, width=96, n_id=1)
    lat, cond, traj, idc, mask = pipeline_oracle.prepare_latents(vsd, vcfg, inp["image"], inp["traj_tensor"],
                                                                 inp["ID_tensor"], 1, 64, 96, 9, inp["latents"])
    assert lat.shape == (1, 16, 3, 4, 6) and torch.equal(lat, inp["latents"])
    assert cond.shape == (1, 16, 1, 4, 6) and idc.shape == (1, 16, 1, 4, 6)
    assert traj.shape == (1, 16, 4, 4, 6) and not traj[:, :, 3].any() and traj[:, :, :3].any()
    assert mask.shape == (1, 1, 3, 4, 6) and not mask[:, :, 0].any() and bool((mask[:, :, 1:] == 1).all())
    mean = torch.tensor(vcfg["latents_mean"]).view(1, 16, 1, 1, 1)
    std = torch.tensor(vcfg["latents_std"]).view(1, 16, 1, 1, 1)
    raw = vae_oracle.encode(vsd, vcfg, inp["image"].unsqueeze(2))[:, :16]
    assert torch.allclose(cond, (raw - mean) / std, atol=1e-6)
    # no ID frame: nothing appended
    out = pipeline_oracle.prepare_latents(vsd, vcfg, inp["image"], inp["traj_tensor"], None, 1, 64, 96, 9, inp["latents"])
    assert out[3] is None and out[2].shape == (1, 16, 3, 4, 6)


def test_oracle_generate_is_loop_plus_decode():
    """``generate`` == prepare_latents -> the reference loop of frameino_b200.sampling over the oracle forward -> final
    first-frame blend -> un-normalise -> decode -> postprocess (two independent spellings of pipeline :809-929)."""
    from frameino_b200.sampling import wan_frameino_denoise
    from oracle import pipeline_oracle, vae_oracle, wan_oracle

    vcfg, vsd, wcfg, wsd = _tiny()
    inp = synth.make_pipeline_inputs(vcfg, 64, num_frames=5, height=64, width=64, n_id=1)
    taps = {}
    video = pipeline_oracle.generate(wsd, wcfg, vsd, vcfg, inp["image"], inp["traj_tensor"], inp["ID_tensor"],
                                     inp["prompt_embeds"], inp["negative_prompt_embeds"], 64, 64, 5,
                                     num_inference_steps=3, guidance_scale=5.0, latents=inp["latents"], taps=taps)
    assert video.shape == (1, 5, 3, 64, 64) and float(video.min()) >= 0 and float(video.max()) <= 1

    lat, cond, traj, idc, mask = pipeline_oracle.prepare_latents(vsd, vcfg, inp["image"], inp["traj_tensor"],
                                                                 inp["ID_tensor"], 1, 64, 64, 5, inp["latents"])

    def tf(hidden_states, timestep, encoder_hidden_states, return_dict=False):
        return (wan_oracle.wan_forward(wsd, wcfg, hidden_states, timestep, encoder_hidden_states),)

    out = wan_frameino_denoise(tf, lat, cond.expand(-1, -1, 2, -1, -1), mask.expand(1, 16, -1, -1, -1), traj, idc,
                               inp["prompt_embeds"], inp["negative_prompt_embeds"], num_steps=3, guidance_scale=5.0,
                               model_dtype=torch.float32)
    out = (1 - mask) * cond + mask * out
    assert torch.allclose(out, taps["latents"], atol=1e-5)
    mean = torch.tensor(vcfg["latents_mean"]).view(1, 16, 1, 1, 1)
    std = torch.tensor(vcfg["latents_std"]).view(1, 16, 1, 1, 1)
    ref = (vae_oracle.decode(vsd, vcfg, out * std + mean) / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4)
    assert torch.allclose(ref, video, atol=1e-4)
    # guidance <= 1: a single forward per step, the negative prompt is not read
    v1 = pipeline_oracle.generate(wsd, wcfg, vsd, vcfg, inp["image"], inp["traj_tensor"], None, inp["prompt_embeds"],
                                  None, 64, 64, 5, num_inference_steps=2, guidance_scale=1.0, latents=inp["latents"],
                                  output_type="latent")
    assert v1.shape == (1, 16, 2, 4, 4) and torch.isfinite(v1).all()


def test_video_processor_round_trip():
    vp = VideoProcessor(16)
    x = torch.rand(3, 32, 48)
    assert torch.equal(vp.preprocess(x, 32, 48), (2 * x - 1)[None])
    y = x * 2 - 1  # already normalised: left alone
    assert torch.equal(vp.preprocess(y, 32, 48), y[None])
    arr = np.random.default_rng(0).random((32, 48, 3), dtype=np.float32)
    assert torch.allclose(vp.preprocess(arr, 32, 48)[0], torch.from_numpy(arr).permute(2, 0, 1) * 2 - 1)
    import PIL.Image

    img = PIL.Image.fromarray((arr * 255).astype(np.uint8))
    p = vp.preprocess(img, 32, 48)
    assert p.shape == (1, 3, 32, 48) and float((p[0] - (torch.from_numpy((arr * 255).astype(np.uint8)).permute(2, 0, 1)
                                                        / 255.0 * 2 - 1)).abs().max()) < 1e-6
    assert vp.preprocess(img, 16, 24).shape == (1, 3, 16, 24)  # PIL inputs are resized
    with pytest.raises(ValueError, match="resize"):
        vp.preprocess(x, 16, 24)
    video = torch.randn(2, 3, 5, 8, 8) * 2
    pt = vp.postprocess_video(video, "pt")
    assert pt.shape == (2, 5, 3, 8, 8) and float(pt.min()) >= 0 and float(pt.max()) <= 1
    assert torch.equal(pt, (video / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4))
    npv = vp.postprocess_video(video, "np")
    assert npv.shape == (2, 5, 8, 8, 3) and npv.dtype == np.float32
    pil = vp.postprocess_video(video, "pil")
    assert len(pil) == 2 and len(pil[0]) == 5 and pil[0][0].size == (8, 8)
    with pytest.raises(ValueError, match="does not exist"):
        vp.postprocess_video(video, "mp4")


class _Dist:
    def __init__(self, m):
        self.m = m

    def mode(self):
        return self.m

    def sample(self, generator=None):
        return self.m + 1


class _Enc:
    def __init__(self, m):
        self.latent_dist = _Dist(m)


def test_retrieve_latents_modes():
    m = torch.zeros(2)
    assert torch.equal(retrieve_latents(_Enc(m), sample_mode="argmax"), m)
    assert torch.equal(retrieve_latents(_Enc(m)), m + 1)
    with pytest.raises(AttributeError):
        retrieve_latents(object())


def _cpu_pipe():
    from frameino_b200.vae import AutoencoderKLWan
    from frameino_b200.wan import WanTransformer3DModel

    vae = AutoencoderKLWan(**synth.with_latent_stats(synth.VAE_TINY))
    tf = WanTransformer3DModel(**synth.WAN_TINY)
    return WanFrameINOPipeline(vae=vae, transformer=tf)


def test_pipeline_validation_errors_match_the_reference():
    """check_inputs (:339-397) and the constructor's scope limits."""
    pipe = _cpu_pipe()
    assert pipe.vae_scale_factor_temporal == 4 and pipe.vae_scale_factor_spatial == 16
    assert pipe.config.expand_timesteps is True and pipe.config.boundary_ratio is None
    img = torch.zeros(1, 3, 64, 64)
    emb = torch.zeros(1, 16, 64)
    with pytest.raises(ValueError, match="divisible by 16"):
        pipe(image=img, prompt_embeds=emb, traj_tensor=torch.zeros(5, 3, 64, 64), height=72, width=64)
    with pytest.raises(ValueError, match="Provide either `prompt` or `prompt_embeds`"):
        pipe(image=img, traj_tensor=torch.zeros(5, 3, 64, 64), height=64, width=64)
    with pytest.raises(ValueError, match="Cannot forward both `prompt`"):
        pipe(image=img, prompt="a", prompt_embeds=emb, height=64, width=64)
    with pytest.raises(ValueError, match="Cannot forward both `negative_prompt`"):
        pipe(image=img, prompt_embeds=emb, negative_prompt="", negative_prompt_embeds=emb, height=64, width=64)
    with pytest.raises(ValueError, match="image_embeds"):
        pipe(image=None, prompt_embeds=emb, height=64, width=64)
    with pytest.raises(ValueError, match="guidance_scale_2"):
        pipe(image=img, prompt_embeds=emb, guidance_scale_2=3.0, height=64, width=64)
    with pytest.raises(ValueError, match="callback_on_step_end_tensor_inputs"):
        pipe(image=img, prompt_embeds=emb, height=64, width=64, callback_on_step_end_tensor_inputs=["nope"])
    with pytest.raises(ValueError, match="traj_tensor"):
        pipe(image=img, prompt_embeds=emb, height=64, width=64)
    # no CPU path, no silent text encoder
    with pytest.raises(RuntimeError, match="no CPU path"):
        pipe(image=img, prompt_embeds=emb, negative_prompt_embeds=emb, traj_tensor=torch.zeros(5, 3, 64, 64),
             height=64, width=64, num_frames=5)
    with pytest.raises(NotImplementedError, match="UMT5"):
        pipe.encode_prompt("a prompt")
    with pytest.raises(NotImplementedError, match="expand_timesteps"):
        WanFrameINOPipeline(vae=pipe.vae, transformer=pipe.transformer, expand_timesteps=False)
    with pytest.raises(NotImplementedError, match="boundary_ratio"):
        WanFrameINOPipeline(vae=pipe.vae, transformer=pipe.transformer, boundary_ratio=0.9)


def test_pipeline_text_encoder_hook_and_scheduler_shift():
    calls = []

    def enc(prompts, max_len):
        calls.append((tuple(prompts), max_len))
        return torch.ones(len(prompts), 4, 64) * len(calls)

    class Sched:
        config = dict(shift=3.0)

    base = _cpu_pipe()
    pipe = WanFrameINOPipeline(vae=base.vae, transformer=base.transformer, text_encoder=enc, scheduler=Sched())
    assert pipe.shift == 3.0
    pos, neg = pipe.encode_prompt(["a", "b"], num_videos_per_prompt=2, max_sequence_length=77)
    assert pos.shape == (4, 4, 64) and neg.shape == (4, 4, 64)
    assert calls == [(("a", "b"), 77), (("", ""), 77)]  # the empty negative prompt of app.py:707
    with pytest.raises(NotImplementedError, match="shift"):
        WanFrameINOPipeline(vae=base.vae, transformer=base.transformer, scheduler=object())


def test_prepare_latents_argument_errors_come_before_any_device_work():
    """pipeline :419-424 (generator list length) and the shape the call needs; Wan2.1 inputs raise."""
    pipe = _cpu_pipe()
    img = torch.zeros(1, 3, 64, 64)
    traj = torch.zeros(5, 3, 64, 64)
    gens = [torch.Generator().manual_seed(i) for i in range(3)]
    with pytest.raises(ValueError, match="list of generators of length 3"):
        pipe.prepare_latents(img, traj, None, 2, 16, 64, 64, 5, torch.float32, torch.device("cpu"), gens, None)
    with pytest.raises(ValueError, match="latents have shape"):
        pipe.prepare_latents(img, traj, None, 1, 16, 64, 64, 5, torch.float32, torch.device("cpu"), None,
                             torch.zeros(1, 16, 3, 4, 4))
    with pytest.raises(NotImplementedError, match="last_image"):
        pipe.prepare_latents(img, traj, None, 1, 16, 64, 64, 5, torch.float32, torch.device("cpu"), None, None, img)
    # the VAE itself refuses CPU tensors (no CPU path), after the argument checks passed
    with pytest.raises(RuntimeError, match="no CPU path"):
        pipe.prepare_latents(img, traj, None, 1, 16, 64, 64, 5, torch.float32, torch.device("cpu"), None,
                             torch.zeros(1, 16, 2, 4, 4))
