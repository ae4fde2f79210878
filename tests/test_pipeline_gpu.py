"""-m gpu: the Wan FrameINO pipeline end to end on the device (frameino_b200.pipeline: 3 VAE encodes -> fused sampler
loop -> VAE decode) against the CPU restatement of the reference pipeline file (oracle/pipeline_oracle.py), same
pixel-space inputs, same initial noise."""
import numpy as np
import pytest
import torch

from conftest import cosine, rel_err
from frameino_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 2e-2
COS = 0.999
H, W, F = 64, 96, 9


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200.pipeline import WanFrameINOPipeline
    from frameino_b200.vae import AutoencoderKLWan
    from frameino_b200.wan import WanTransformer3DModel
    from oracle import wan_oracle

    vcfg = synth.with_latent_stats(synth.VAE_TINY)
    vsd = synth.make_vae_state_dict(synth.VAE_TINY, seed=1)
    wsd = synth.make_wan_state_dict(synth.WAN_TINY, seed=0, dtype=torch.bfloat16)
    vae = AutoencoderKLWan(**vcfg)
    vae.load_state_dict(vsd, strict=True)
    vae = vae.cuda().eval()
    tf = WanTransformer3DModel(**synth.WAN_TINY)
    tf.load_state_dict(wsd, strict=True)
    tf = tf.to_inference_dtype(torch.bfloat16).cuda().eval()
    pipe = WanFrameINOPipeline(vae=vae, transformer=tf)
    inp = synth.make_pipeline_inputs(vcfg, 64, num_frames=F, height=H, width=W, n_id=1)
    return pipe, inp, vcfg, vsd, wan_oracle.WanConfig(**synth.WAN_TINY), wsd


def _call(pipe, inp, **kw):
    args = dict(image=inp["image"], traj_tensor=inp["traj_tensor"], ID_tensor=inp["ID_tensor"],
                prompt_embeds=inp["prompt_embeds"].cuda(), negative_prompt_embeds=inp["negative_prompt_embeds"].cuda(),
                latents=inp["latents"], height=H, width=W, num_frames=F, num_inference_steps=4, guidance_scale=5.0,
                output_type="pt")
    args.update(kw)
    return pipe(**args)


def test_prepare_latents_vs_oracle(setup):
    """pipeline :400-536 — the three VAE encodes, the normalisation, the ID padding, the mask."""
    from oracle import pipeline_oracle

    pipe, inp, vcfg, vsd, _, _ = setup
    want = pipeline_oracle.prepare_latents(vsd, vcfg, inp["image"], inp["traj_tensor"], inp["ID_tensor"], 1, H, W, F,
                                           inp["latents"])
    got = pipe.prepare_latents(inp["image"], inp["traj_tensor"], inp["ID_tensor"], 1, 16, H, W, F, torch.float32,
                               torch.device("cuda"), None, inp["latents"])
    names = ["latents", "latent_condition", "traj_latents", "ID_latent_condition", "first_frame_mask"]
    for n, g, w_ in zip(names, got, want):
        assert g.shape == w_.shape and g.dtype == torch.float32, n
        if n in ("latents", "first_frame_mask"):
            assert torch.equal(g.cpu(), w_), n
        else:
            assert rel_err(g, w_) <= TOL, (n, rel_err(g, w_))
            assert cosine(g, w_) >= COS, n
    assert not got[2][:, :, -1].any()  # zero trajectory latents on the ID frame (:517-518)
    # no ID image: four values + None
    got = pipe.prepare_latents(inp["image"], inp["traj_tensor"], None, 1, 16, H, W, F, torch.float32,
                               torch.device("cuda"), None, inp["latents"])
    assert got[3] is None and got[2].shape[2] == 3
    empty = inp["ID_tensor"][:, :, :0]
    assert pipe.prepare_latents(inp["image"], inp["traj_tensor"], empty, 1, 16, H, W, F, torch.float32,
                                torch.device("cuda"), None, inp["latents"])[3] is None  # :489


def test_pipeline_end_to_end_vs_oracle(setup):
    """pixels -> pixels: final latents cosine >= 0.999 (north_star), decoded video close."""
    from oracle import pipeline_oracle

    pipe, inp, vcfg, vsd, wcfg, wsd = setup
    taps = {}
    wsd32 = {k: v.float() for k, v in wsd.items()}
    want = pipeline_oracle.generate(wsd32, wcfg, vsd, vcfg, inp["image"], inp["traj_tensor"], inp["ID_tensor"],
                                    inp["prompt_embeds"].bfloat16().float(),
                                    inp["negative_prompt_embeds"].bfloat16().float(), H, W, F, num_inference_steps=4,
                                    guidance_scale=5.0, latents=inp["latents"], taps=taps)
    lat = _call(pipe, inp, output_type="latent").frames
    assert lat.shape == (1, 16, 3, 4, 6) and lat.dtype == torch.float32
    assert cosine(lat, taps["latents"]) >= COS, cosine(lat, taps["latents"])
    # frame 0 of the result is the clean first-frame condition (:914-915)
    assert rel_err(lat[:, :, 0], taps["latents"][:, :, 0]) <= TOL
    video = _call(pipe, inp).frames
    assert video.shape == (1, F, 3, H, W) == want.shape
    assert float(video.min()) >= 0 and float(video.max()) <= 1
    err = (video.float().cpu() - want).abs()
    assert float(err.mean()) <= 2e-2, float(err.mean())
    assert cosine(video.float() - 0.5, want - 0.5) >= 0.99


def test_pipeline_vs_the_reference_pipeline_files_own_output(setup, golden_dir):
    """The native call against tests/golden/pipeline_golden.pt — outputs of the reference's own pipeline file executed
    on CPU in fp32 around the reference transformer and VAE (make_golden.py pipeline)."""
    import os

    pipe, inp, *_ = setup
    gold = torch.load(os.path.join(golden_dir, "pipeline_golden.pt"))
    got = pipe.prepare_latents(inp["image"], inp["traj_tensor"], inp["ID_tensor"], 1, 16, H, W, F, torch.float32,
                               torch.device("cuda"), None, inp["latents"])
    for n, g in zip(["latents", "latent_condition", "traj_latents", "ID_latent_condition", "first_frame_mask"], got):
        assert rel_err(g, gold["prepare." + n]) <= TOL and cosine(g, gold["prepare." + n]) >= COS, n
    lat = _call(pipe, inp, output_type="latent").frames
    assert cosine(lat, gold["call.latents"]) >= COS, cosine(lat, gold["call.latents"])
    video = _call(pipe, inp).frames
    assert float((video.float().cpu() - gold["call.video"]).abs().mean()) <= 2e-2
    lat1 = _call(pipe, inp, ID_tensor=inp["ID_tensor"][:, :, :0], guidance_scale=1.0, negative_prompt_embeds=None,
                 output_type="latent").frames
    assert cosine(lat1, gold["call_noid_nocfg.latents"]) >= COS


def test_fused_and_plain_loop_agree_bit_for_bit(setup):
    pipe, inp, *_ = setup
    a = _call(pipe, inp, output_type="latent").frames
    b = _call(pipe, inp, output_type="latent", fused=False).frames
    assert torch.equal(a, b)


def test_output_types_callbacks_and_generators(setup):
    pipe, inp, *_ = setup
    out = _call(pipe, inp, output_type="np", return_dict=False)
    assert isinstance(out, tuple) and isinstance(out[0], np.ndarray) and out[0].shape == (1, F, H, W, 3)
    ref = _call(pipe, inp).frames
    assert np.array_equal(out[0], ref.permute(0, 1, 3, 4, 2).float().cpu().numpy())
    # callback: called once per step with the latents; returning latents replaces them (:893-899)
    seen = []

    def cb(p, i, t, kw):
        seen.append((i, float(t), tuple(kw["latents"].shape)))
        return {"latents": torch.zeros_like(kw["latents"])} if i == 3 else {}

    lat = _call(pipe, inp, output_type="latent", callback_on_step_end=cb).frames
    assert [s[0] for s in seen] == [0, 1, 2, 3] and seen[0][2] == (1, 16, 3, 4, 6)
    assert seen[0][1] > seen[-1][1] > 0  # timesteps descend from ~1000
    assert not lat[:, :, 1:].any() and lat[:, :, 0].any()  # zeroed by the callback, frame 0 re-blended from the condition
    assert pipe.num_timesteps == 4 and pipe.guidance_scale == 5.0 and pipe.do_classifier_free_guidance
    # same for the plain loop
    seen.clear()
    lat2 = _call(pipe, inp, output_type="latent", callback_on_step_end=cb, fused=False).frames
    assert torch.equal(lat, lat2) and len(seen) == 4
    # initial noise from a generator: CPU generator -> the CPU stream's numbers, reproducible
    g = torch.Generator().manual_seed(3)
    a = _call(pipe, inp, latents=None, generator=g, output_type="latent").frames
    want0 = torch.randn(1, 16, 3, 4, 6, generator=torch.Generator().manual_seed(3))
    b = _call(pipe, inp, latents=want0, output_type="latent").frames
    assert torch.equal(a, b)
    # guidance <= 1: one forward per step, no negative prompt needed
    c = _call(pipe, inp, guidance_scale=1.0, negative_prompt_embeds=None, output_type="latent").frames
    assert torch.isfinite(c).all() and not pipe.do_classifier_free_guidance
    # CFG without a negative prompt and without a text encoder: loud
    with pytest.raises(NotImplementedError, match="UMT5"):
        _call(pipe, inp, negative_prompt_embeds=None)
    # no ID image: nothing appended to the sequence
    e = _call(pipe, inp, ID_tensor=None, output_type="latent").frames
    assert e.shape == (1, 16, 3, 4, 6) and torch.isfinite(e).all()
    assert torch.equal(e, _call(pipe, inp, ID_tensor=None, output_type="latent", fused=False).frames)
    # num_frames is rounded down to 1 + 4k like the reference (:707-711)
    d = _call(pipe, inp, num_frames=F + 2, output_type="latent").frames
    assert d.shape[2] == 3
