"""-m gpu: the device-side denoise-loop glue — the two fused kernels against the reference's
tensor-op chain (bit exact), the per-prompt text state (exact, cached, invalidated), and the fused sampler loop against
the plain loop (same native model) and against the CPU oracle loop."""
import pytest
import torch

from conftest import cosine
from frameino_b200 import synth

pytestmark = pytest.mark.gpu
COS_TOL = 0.999


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _native(cfg, sd):
    from frameino_b200.wan import WanTransformer3DModel

    m = WanTransformer3DModel(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.to_inference_dtype(torch.bfloat16).cuda().eval()


def _loop_inputs(b, c, f, h, w, n_id, seed=7, mask_channels=1):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(b, c, f, h, w, generator=g)
    cond = torch.zeros(b, c, f, h, w)
    cond[:, :, 0] = torch.randn(b, c, h, w, generator=g)
    mask = torch.ones(1, mask_channels, f, h, w)
    mask[:, :, 0] = 0
    traj = torch.randn(b, c, f + n_id, h, w, generator=g)
    traj[:, :, f:] = 0
    idl = torch.randn(b, c, n_id, h, w, generator=g) if n_id else None
    return lat, cond, mask, traj, idl


@pytest.mark.parametrize("b,c,f,h,w,n_id", [(1, 16, 3, 16, 16, 1), (2, 4, 5, 6, 10, 2), (1, 48, 2, 44, 80, 0),
                                            (1, 3, 1, 2, 2, 1)])
def test_pack_model_input_is_the_reference_chain_bit_for_bit(b, c, f, h, w, n_id):
    """pipeline_wan_i2v_motion_FrameINO.py:829-858 + the patchify of transformer_wan.py:486-487."""
    from frameino_b200 import ops

    lat, cond, mask, traj, idl = _loop_inputs(b, c, f, h, w, n_id)
    mask[0, 0, 1:, ::3] = 0.25  # the blend is exercised with a non-binary mask too
    x = (1 - mask) * cond + mask * lat
    if idl is not None:
        x = torch.cat([x, idl], dim=2)
    x = torch.cat([x, traj], dim=1).to(torch.bfloat16)
    ft = f + n_id
    # Conv3d(k = s = (1,2,2)) input rows: [(b, f, h/2, w/2), (c, 1, 2, 2)]
    want = x.view(b, 2 * c, ft, 1, h // 2, 2, w // 2, 2).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(-1, 2 * c * 4)
    got = ops.wan_pack_model_input(lat.cuda(), cond.cuda(), mask[0, 0].contiguous().cuda(),
                                   None if idl is None else idl.cuda(), traj.cuda(), (1, 2, 2))
    assert got.dtype == torch.bfloat16 and got.shape == want.shape
    assert torch.equal(got.cpu(), want)
    # and it is what the model's own patchify makes of the materialised 5-D input
    xs = x.cuda()
    rows = ops.patchify(xs, tuple(xs.shape), xs.stride(), (1, 2, 2))
    assert torch.equal(rows, got)


@pytest.mark.parametrize("b,c,f,h,w,n_id", [(1, 16, 3, 16, 16, 1), (2, 4, 5, 6, 10, 2), (1, 48, 2, 44, 80, 0)])
@pytest.mark.parametrize("cfg_on", [True, False])
def test_cfg_euler_step_is_the_reference_chain_bit_for_bit(b, c, f, h, w, n_id, cfg_on):
    """pipeline :882 (guidance), :886 (ID-frame drop), :891 (Euler) on the un-patchified outputs (transformer_wan.py
    :539-543)."""
    from frameino_b200 import ops

    g = torch.Generator().manual_seed(3)
    ft = f + n_id
    tokens = ft * (h // 2) * (w // 2)
    y_c = torch.randn(b * tokens, 4 * c, generator=g).bfloat16()
    y_u = torch.randn(b * tokens, 4 * c, generator=g).bfloat16()
    lat = torch.randn(b, c, f, h, w, generator=g)
    guidance, dsigma = 4.3, -0.0371

    def unpatch(y):  # :539-543
        t = y.view(b, ft, h // 2, w // 2, 1, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return t.reshape(b, c, ft, h, w)

    v = unpatch(y_c)
    if cfg_on:
        vu = unpatch(y_u)
        v = (vu + guidance * (v - vu)).float()  # :882 on the bf16 outputs
    v = v[:, :, :f].float()
    want = lat + torch.tensor(dsigma, dtype=torch.float32) * v
    got = ops.wan_cfg_euler_step(lat.cuda().clone(), y_c.cuda(), y_u.cuda() if cfg_on else None, n_id, (1, 2, 2),
                                 guidance, float(torch.tensor(dsigma, dtype=torch.float32)))
    assert torch.equal(got.cpu(), want)


def test_loop_kernels_reject_bad_arguments():
    from frameino_b200 import _lib, ops

    lat = torch.zeros(1, 4, 2, 4, 4, device="cuda")
    with pytest.raises(ValueError):
        ops.wan_pack_model_input(lat, lat, torch.zeros(3, 4, 4, device="cuda"), None, lat, (1, 2, 2))
    with pytest.raises(_lib.FinoError):  # H not divisible by the patch
        odd = torch.zeros(1, 4, 2, 3, 4, device="cuda")
        ops.wan_pack_model_input(odd, odd, torch.zeros(2, 3, 4, device="cuda"), None, odd, (1, 2, 2))


def test_text_state_cache_is_exact_and_invalidates():
    from frameino_b200 import ops

    cfg = synth.WAN_SMALL
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    hidden, ts, text = synth.make_wan_inputs(cfg, 3, 16, 16, n_id=1, text_len=16, text_true_len=11,
                                             dtype=torch.bfloat16)
    model = _native(cfg, sd)
    hidden, ts, text = hidden.cuda(), ts.cuda(), text.cuda()

    def fwd():
        n0 = ops.launch_count()
        out = model(hidden_states=hidden, timestep=ts, encoder_hidden_states=text, return_dict=False)[0]
        return out, ops.launch_count() - n0

    base, n_plain = fwd()
    with model.cache_context("cond"):
        first, n_first = fwd()
        second, n_second = fwd()
    assert torch.equal(base, first) and torch.equal(base, second)
    layers = cfg["num_layers"]
    assert n_first == n_plain
    assert n_second == n_plain - 2 - 2 * layers  # text MLP (2 GEMMs) + per layer K/V GEMM and key norm

    # a different context name has its own state; the same name with another prompt recomputes
    text2 = (text.float() * 0.5).bfloat16()
    with model.cache_context("uncond"):
        other = model(hidden_states=hidden, timestep=ts, encoder_hidden_states=text2, return_dict=False)[0]
    assert not torch.equal(other, base)
    with model.cache_context("cond"):
        again, n_again = fwd()
        assert n_again == n_second and torch.equal(again, base)
        swapped = model(hidden_states=hidden, timestep=ts, encoder_hidden_states=text2, return_dict=False)[0]
    assert torch.equal(swapped, other)

    # in-place edits of the prompt or of a text-side weight invalidate the state
    with model.cache_context("cond"):
        fwd()
        text.mul_(0.5)
        edited, n_edit = fwd()
        assert n_edit == n_plain
        with torch.no_grad():
            model.blocks[0].attn2.to_v.weight.mul_(2.0)
        reweighted, n_rw = fwd()
        assert n_rw == n_plain
    assert not torch.equal(edited, base) and not torch.equal(reweighted, edited)
    model.clear_text_cache()
    ref, _ = fwd()
    assert torch.equal(ref, reweighted)


@pytest.mark.parametrize("n_id,do_cfg", [(1, True), (0, True), (1, False)])
def test_fused_loop_equals_plain_loop(n_id, do_cfg):
    """Same native model, same inputs: the fused glue changes no arithmetic, so the final latents are identical."""
    from frameino_b200.sampling import wan_frameino_denoise, wan_frameino_denoise_fused

    cfg = synth.WAN_TINY
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    model = _native(cfg, sd)
    lat, cond, mask, traj, idl = _loop_inputs(1, 16, 3, 16, 16, n_id)
    g = torch.Generator().manual_seed(5)
    pos = torch.randn(1, 16, 64, generator=g).bfloat16().cuda()
    neg = torch.zeros(1, 16, 64).bfloat16().cuda() if do_cfg else None
    if idl is None:
        idl_plain = torch.zeros(1, 16, 0, 16, 16)
    else:
        idl_plain = idl
    plain = wan_frameino_denoise(model, lat.cuda(), cond.cuda(), mask.expand(1, 16, -1, -1, -1).cuda(), traj.cuda(),
                                 idl_plain.cuda(), pos, neg, num_steps=6)
    fused = wan_frameino_denoise_fused(model, lat.cuda(), cond.cuda(), mask.cuda(), traj.cuda(),
                                       None if idl is None else idl.cuda(), pos, neg, num_steps=6)
    assert torch.isfinite(fused).all()
    assert torch.equal(plain, fused)


@pytest.mark.parametrize("num_steps", [10, 50])
def test_fused_loop_final_latent_cosine_vs_oracle(num_steps):
    """Config 4 at test size through the fused loop, up to the reference's full 50-step schedule
    (pipeline_wan_i2v_motion_FrameINO.py:809-908): final latent cosine >= 0.999 vs the CPU oracle loop."""
    from frameino_b200.sampling import wan_frameino_denoise, wan_frameino_denoise_fused
    from oracle import wan_oracle

    cfg = synth.WAN_TINY
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    lat, cond, mask, traj, idl = _loop_inputs(1, 16, 3, 16, 16, 1)
    pos = torch.randn(1, 16, 64, generator=torch.Generator().manual_seed(5)).bfloat16()
    neg = torch.zeros(1, 16, 64).bfloat16()
    ocfg = wan_oracle.WanConfig(**cfg)

    def oracle_tf(hidden_states, timestep, encoder_hidden_states, return_dict=False):
        return (wan_oracle.wan_forward(sd, ocfg, hidden_states, timestep, encoder_hidden_states),)

    ref = wan_frameino_denoise(oracle_tf, lat, cond, mask.expand(1, 16, -1, -1, -1), traj, idl, pos, neg,
                               num_steps=num_steps)
    model = _native(cfg, sd)
    out = wan_frameino_denoise_fused(model, lat.cuda(), cond.cuda(), mask.cuda(), traj.cuda(), idl.cuda(), pos.cuda(),
                                     neg.cuda(), num_steps=num_steps)
    assert cosine(out, ref) >= COS_TOL


def test_fused_loop_rejects_non_binary_mask():
    from frameino_b200.sampling import wan_frameino_denoise_fused

    cfg = synth.WAN_TINY
    model = _native(cfg, synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16))
    lat, cond, mask, traj, idl = _loop_inputs(1, 16, 3, 16, 16, 1)
    mask[0, 0, 1, 0, 0] = 0.5
    pos = torch.zeros(1, 16, 64).bfloat16().cuda()
    with pytest.raises(NotImplementedError, match="0/1"):
        wan_frameino_denoise_fused(model, lat.cuda(), cond.cuda(), mask.cuda(), traj.cuda(), idl.cuda(), pos, None,
                                   num_steps=1)
