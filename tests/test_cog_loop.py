"""CogVideoX FrameINO sampler-loop glue (frameino_b200.sampling.cog_frameino_denoise): the pieces on CPU, the native
model through the loop against the CPU oracle on the GPU."""
import math

import pytest
import torch

from frameino_b200 import synth
from frameino_b200.sampling import (cog_frameino_denoise, ddim_v_step, dynamic_cfg_scale,
                                    scaled_linear_alphas_cumprod)


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
