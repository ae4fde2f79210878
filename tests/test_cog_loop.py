"""CogVideoX FrameINO sampler-loop glue (frameino_b200.sampling.cog_frameino_denoise): the pieces on CPU, the native
model through the loop against the CPU oracle on the GPU."""
import math

import pytest
import torch

from frameino_b200 import synth
from frameino_b200.sampling import (CogVideoXDPMSchedule, cog_frameino_denoise, ddim_v_step, dynamic_cfg_scale,
                                    rescale_zero_terminal_snr, scaled_linear_alphas_cumprod)


def test_dynamic_cfg_schedule_is_the_pipeline_formula():
    # pipeline_cogvideox_i2v_motion_FrameINO.py:906-909
    for steps, t, gs in [(50, 999.0, 6.0), (50, 500.0, 6.0), (50, 0.0, 6.0), (20, 10.0, 3.5)]:
        want = 1 + gs * ((1 - math.cos(math.pi * ((steps - t) / steps) ** 5.0)) / 2)
        assert dynamic_cfg_scale(gs, steps, t) == want


def test_ddim_v_step_identities():
    g = torch.Generator().manual_seed(0)
    x, v = torch.randn(2, 3, 4, generator=g), torch.randn(2, 3, 4, generator=g)
    assert torch.allclose(ddim_v_step(v, x, 0.37, 0.37), x, atol=1e-6)           # no change of noise level: identity
    a = 0.6
    x0 = a ** 0.5 * x - (1 - a) ** 0.5 * v
    assert torch.allclose(ddim_v_step(v, x, a, 1.0), x0, atol=1e-6)              # last step lands on the x0 estimate
    ac = scaled_linear_alphas_cumprod()
    assert ac.shape == (1000,) and bool((ac[:-1] > ac[1:]).all()) and 0 < float(ac[-1]) < float(ac[0]) < 1


def test_dpm_schedule_tables():
    sch = CogVideoXDPMSchedule()
    ac = sch.alphas_cumprod
    assert float(ac[-1]) == 0.0 and abs(float(ac[0]) - float(scaled_linear_alphas_cumprod()[0])) < 1e-6  # zero terminal SNR
    assert bool((ac[:-1] > ac[1:]).all())
    ts = sch.set_timesteps(50)
    assert ts[0] == 999 and ts[-1] == 19 and len(ts) == 50 and bool((ts[:-1] - ts[1:] == 20).all())  # trailing spacing
    assert list(CogVideoXDPMSchedule().set_timesteps(3)) == [999, 666, 332]
    r = rescale_zero_terminal_snr(torch.tensor([0.9, 0.5, 0.1]))
    assert abs(float(r[0]) - 0.9) < 1e-6 and float(r[-1]) == 0.0


def test_dpm_step_is_dpm_solver_pp_2m_sde():
    """``CogVideoXDPMSchedule.step`` against the closed form of DPM-Solver++(2M) SDE written with
    alpha = sqrt(a_bar), sigma = sqrt(1 - a_bar), lambda = log(alpha / sigma), h = lambda_prev - lambda_t:
        x_prev = (sigma_prev / sigma_t) e^{-h} x + alpha_prev (1 - e^{-2h}) D + sigma_prev sqrt(1 - e^{-2h}) z
        D = x0_t (first / last step),   D = (1 + 1/(2r)) x0_t - (1/(2r)) x0_back,  r = h_last / h
    in float64, over a whole 10-step trajectory of a linear toy model (v = 0.3 x + c)."""
    g = torch.Generator().manual_seed(0)
    sch = CogVideoXDPMSchedule()
    ts = sch.set_timesteps(10)
    x = torch.randn(2, 3, 5, generator=g)
    xd = x.double()
    c = torch.randn(2, 3, 5, generator=g)
    old = None
    old_d = None
    ab = sch.alphas_cumprod.double()
    for i, t in enumerate(ts):
        z = torch.randn(2, 3, 5, generator=g)
        v = 0.3 * x + c
        x_new, old_new = sch.step(v, old, int(t), int(ts[i - 1]) if i > 0 else None, x, noise=z)
        # closed form (float64)
        vd = 0.3 * xd + c.double()
        prev_t = int(t) - 1000 // 10
        a_t, a_p = ab[int(t)], (ab[prev_t] if prev_t >= 0 else torch.tensor(1.0, dtype=torch.float64))
        al_t, sg_t, al_p, sg_p = a_t.sqrt(), (1 - a_t).sqrt(), a_p.sqrt(), (1 - a_p).sqrt()
        x0 = al_t * xd - sg_t * vd
        h = torch.log(al_p / sg_p) - torch.log(al_t / sg_t)
        d = x0
        if old_d is not None and prev_t >= 0:
            a_b = ab[int(ts[i - 1])]
            h_last = torch.log(al_t / sg_t) - torch.log(a_b.sqrt() / (1 - a_b).sqrt())
            r = h_last / h
            d = (1 + 1 / (2 * r)) * x0 - (1 / (2 * r)) * old_d
        e2h = torch.exp(-2 * h)
        want = (sg_p / sg_t) * torch.exp(-h) * xd + al_p * (1 - e2h) * d + sg_p * torch.sqrt(1 - e2h) * z.double()
        assert torch.isfinite(x_new).all()
        assert torch.allclose(x_new.double(), want, rtol=2e-4, atol=2e-4), (i, float((x_new.double() - want).abs().max()))
        assert torch.allclose(old_new.double(), x0, rtol=2e-4, atol=2e-4)
        x, old, xd, old_d = x_new, old_new, want, x0
    # the last step (prev_timestep < 0) adds no noise and lands on the denoised estimate
    assert torch.allclose(x.double(), old_d, atol=1e-3)


def test_cog_loop_with_the_dpm_scheduler_calls_step_like_the_pipeline():
    """:918-926 — step(noise_pred, old_pred_original_sample, t, timesteps[i-1] if i > 0 else None, latents)."""
    cfg, (lat, img, traj, idl, text) = _case(1)
    calls = []

    class Spy(CogVideoXDPMSchedule):
        def step(self, model_output, old, timestep, timestep_back, sample, **kw):
            calls.append((old is None, timestep, timestep_back))
            return super().step(model_output, old, timestep, timestep_back, sample, **kw)

    def tf(hidden_states, encoder_hidden_states, timestep, image_rotary_emb, return_dict=False):
        return (hidden_states[:, :, :16] * 0.5,)

    sch = Spy()
    ts = sch.set_timesteps(4)
    out = cog_frameino_denoise(tf, lat, img, traj, idl, text, None, ts, guidance_scale=6.0, model_dtype=torch.float32,
                               scheduler=sch, generator=torch.Generator().manual_seed(1))
    assert calls == [(True, 999, None), (False, 749, 999), (False, 499, 749), (False, 249, 499)]
    assert out.shape == lat.shape and torch.isfinite(out).all()


def _case(n_id):
    cfg = synth.COG_TINY
    g = torch.Generator().manual_seed(3)
    f, c, h, w = 3, 16, 12, 16
    lat = torch.randn(1, f, c, h, w, generator=g)
    img = torch.zeros(1, f, c, h, w)
    img[:, 0] = torch.randn(1, c, h, w, generator=g)
    traj = torch.randn(1, f, c, h, w, generator=g)
    idl = torch.randn(1, n_id, c, h, w, generator=g) if n_id else None
    text = torch.randn(2, cfg["max_text_seq_length"], cfg["text_embed_dim"], generator=g)
    return cfg, (lat, img, traj, idl, text)


def test_loop_builds_the_pipeline_inputs_and_steps():
    """A recording transformer checks the tensors the loop hands over (:853-881) and the arithmetic after it."""
    cfg, (lat, img, traj, idl, text) = _case(1)
    seen = []

    def tf(hidden_states, encoder_hidden_states, timestep, image_rotary_emb, return_dict=False):
        seen.append((hidden_states.clone(), timestep.clone()))
        b, f = hidden_states.shape[:2]
        return (hidden_states[:, :, :16] * 0.5,)

    ts = torch.tensor([900, 500, 100])
    ac = scaled_linear_alphas_cumprod()
    out = cog_frameino_denoise(tf, lat, img, traj, idl, text, None, ts, ac, guidance_scale=6.0, model_dtype=torch.float32)
    x0, t0 = seen[0]
    assert x0.shape == (2, 4, 48, 12, 16) and torch.equal(t0, torch.tensor([900, 900]))
    assert torch.equal(x0[0], x0[1])                                              # CFG batch = two copies (:853)
    assert torch.equal(x0[0, :3, :16], lat[0]) and torch.equal(x0[0, 3, :16], idl[0, 0])
    assert torch.equal(x0[0, :3, 16:32], img[0]) and float(x0[0, 3, 16:].abs().max()) == 0.0   # zero padding (:869-873)
    assert torch.equal(x0[0, :3, 32:], traj[0])
    # both CFG halves return the same v, so guidance changes nothing and step 1 is a plain DDIM step on v = 0.5 x
    want = ddim_v_step(0.5 * lat, lat, ac[900], ac[500])
    assert torch.allclose(seen[1][0][0, :3, :16], want, atol=1e-6)
    assert out.shape == lat.shape and out.dtype == text.dtype and len(seen) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("dynamic", [False, True])
def test_cog_loop_native_vs_oracle(dynamic):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from conftest import cosine
    from frameino_b200.cogvideox import CogVideoXTransformer3DModel
    from oracle import cog_oracle

    cfg, (lat, img, traj, idl, text) = _case(1)
    sd = synth.make_cog_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    cos, sin = cog_oracle.cog_rope_3d(64, (6, 8), 3, 1)
    text = text.bfloat16()

    def oracle_tf(hidden_states, encoder_hidden_states, timestep, image_rotary_emb, return_dict=False):
        return (cog_oracle.cog_forward(sd, cfg, hidden_states, encoder_hidden_states, timestep, image_rotary_emb),)

    ts = torch.linspace(999, 0, 6).long()
    ac = scaled_linear_alphas_cumprod()
    ref = cog_frameino_denoise(oracle_tf, lat, img, traj, idl, text, (cos, sin), ts, ac, use_dynamic_cfg=dynamic)
    model = CogVideoXTransformer3DModel(**cfg)
    model.load_state_dict(sd)
    model = model.to_inference_dtype(torch.bfloat16).cuda().eval()
    out = cog_frameino_denoise(model, lat.cuda(), img.cuda(), traj.cuda(), idl.cuda(), text.cuda(),
                               (cos.cuda(), sin.cuda()), ts, ac, use_dynamic_cfg=dynamic)
    assert torch.isfinite(out).all()
    assert cosine(out, ref) >= 0.999
