"""-m gpu: the Wan VAE path — the implicit-GEMM convolution and the channels-last helper kernels
against torch on the CPU, and the native AutoencoderKLWan (encode + decode, the reference's chunking and caches) against
the whole-sequence CPU oracle and the golden outputs produced by the reference's own source file."""
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import cosine, rel_err
from frameino_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 2e-2
COS = 0.999


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from frameino_b200 import ops as _ops

    return _ops


def _pack(w):  # [Co, Ci, kt, kh, kw] -> [up8(Co), taps * ceil(Ci/64)*64] bf16, tap-major
    co, ci = w.shape[:2]
    taps = w[0, 0].numel()
    cinp = (ci + 63) // 64 * 64
    wp = torch.zeros((co + 7) // 8 * 8, taps, cinp)
    wp[:co, :, :ci] = w.reshape(co, ci, taps).permute(0, 2, 1)
    return wp.reshape(wp.shape[0], -1).bfloat16()


@pytest.mark.parametrize("t,h,w,ci,co,k,stride,st", [
    (3, 9, 21, 24, 40, (3, 3, 3), 1, 1),      # ragged tile edges, channel padding (24 -> 64), N tile 128
    (2, 16, 32, 128, 264, (3, 3, 3), 1, 1),   # two K chunks per tap, N tile 256 with a ragged second tile
    (1, 8, 16, 64, 16, (1, 3, 3), 1, 1),      # 2-D conv, narrow N tile (32)
    (4, 12, 20, 72, 72, (3, 1, 1), 1, 1),     # time_conv
    (5, 10, 14, 64, 64, (3, 1, 1), 1, 2),     # stride-2 time_conv (downsample3d)
    (2, 18, 34, 32, 32, (1, 3, 3), 2, 1),     # stride-2 Conv2d behind ZeroPad2d((0,1,0,1)) (downsample2d)
    (2, 10, 18, 160, 320, (3, 3, 3), 1, 1),   # encoder widths: 2.5 K chunks per tap (OOB channel fill), N tile 160 x 2
    (1, 8, 16, 320, 160, (3, 3, 3), 1, 1),    # N tile 160, one tile
])
def test_conv3d_cl_matches_torch(ops, t, h, w, ci, co, k, stride, st):
    g = torch.Generator().manual_seed(0)
    kt, kh, kw = k
    x = torch.randn(t + (kt - 1 if st == 1 else 0), h, w, ci, generator=g).bfloat16()  # history frames included
    wt = (torch.randn(co, ci, kt, kh, kw, generator=g) / (ci * kt * kh * kw) ** 0.5).bfloat16()
    b = (torch.randn(co, generator=g) * 0.1).bfloat16()
    xin = x.float().permute(3, 0, 1, 2)[None]  # [1, C, T, H, W]
    if stride == 2:
        ref = F.conv3d(F.pad(xin, (0, 1, 0, 1)), wt.float(), b.float(), stride=(1, 2, 2))
        pad, out_hw = (0, 0), (h // 2, w // 2)
    else:
        ref = F.conv3d(F.pad(xin, (kw // 2, kw // 2, kh // 2, kh // 2)), wt.float(), b.float(), stride=(st, 1, 1))
        pad, out_hw = (kh // 2, kw // 2), None
    ref = ref[0].permute(1, 2, 3, 0)  # [T_out, H, W, Co]
    bp = torch.zeros((co + 7) // 8 * 8).bfloat16()
    bp[:co] = b
    y = ops.conv3d_cl(x.cuda(), _pack(wt.float()).cuda(), bp.cuda(), k, pad_hw=pad, stride_hw=stride, stride_t=st,
                      out_hw=out_hw)
    assert y.shape[:3] == ref.shape[:3]
    assert rel_err(y[..., :co], ref) <= 1e-2
    if y.shape[-1] > co:
        assert float(y[..., co:].float().abs().max()) == 0.0
    # residual epilogue + strided (frame-interleaved) output
    if stride == 1 and st == 1 and co % 8 == 0:
        res = torch.randn(ref.shape, generator=g).bfloat16()
        big = torch.zeros(2 * ref.shape[0], *ref.shape[1:], dtype=torch.bfloat16, device="cuda")
        rbig = torch.zeros_like(big)
        rbig[1::2] = res.cuda()
        ops.conv3d_cl(x.cuda(), _pack(wt.float()).cuda(), bp.cuda(), k, pad_hw=pad, out=big[1::2], residual=rbig[1::2])
        assert rel_err(big[1::2], ref + res.float()) <= 1e-2
        assert float(big[0::2].float().abs().max()) == 0.0


def test_vae_row_kernels_match_torch(ops):
    g = torch.Generator().manual_seed(1)
    for c in (16, 64, 160, 256, 320, 640, 1024):
        x = torch.randn(37, c, generator=g).bfloat16()
        gamma = 1 + 0.1 * torch.randn(c, generator=g)
        ref = F.normalize(x.float(), dim=1) * c ** 0.5 * gamma
        assert rel_err(ops.rms_act_cl(x.cuda(), gamma.cuda(), silu=False), ref) <= 1e-2
        assert rel_err(ops.rms_act_cl(x.cuda(), gamma.cuda(), silu=True), F.silu(ref)) <= 1e-2
    x = torch.randn(2, 3, 5, 16, generator=g).bfloat16()
    up = ops.upsample2x_cl(x.cuda()).cpu()
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest-exact").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    s = torch.randn(50, 3520, generator=g) * 30
    p = torch.full((50, 3528), 7.0).bfloat16().cuda()
    ops.softmax_rows(s.cuda(), 1024 ** -0.5, p)
    assert rel_err(p[:, :3520], torch.softmax(s * 1024 ** -0.5, dim=1)) <= 1e-2
    assert float(p[:, 3520:].float().abs().max()) == 0.0


def test_vae_shortcut_and_layout_kernels_match_the_oracle(ops):
    from oracle import vae_oracle

    g = torch.Generator().manual_seed(2)
    # DupUp3D: (c_in, c_out, ft, fs) of the three up blocks of the Wan2.2 decoder, scaled down
    for ci, co, ft, fs, first in [(32, 32, 2, 2, True), (32, 32, 2, 2, False), (32, 16, 1, 2, False)]:
        src = torch.randn(1, ci, 3, 4, 6, generator=g).bfloat16()
        ref = vae_oracle.dup_up3d(src.float(), co, ft, fs)
        if not first:  # later chunks keep every duplicated frame
            ref = torch.cat([ref[:, :, :1], ref], dim=2) if ft == 2 else ref
        y0 = torch.randn(ref.shape, generator=g).bfloat16()
        y = y0.permute(0, 2, 3, 4, 1)[0].contiguous().cuda()
        ops.dupup_add_cl(y, src.permute(0, 2, 3, 4, 1)[0].contiguous().cuda(), ft, fs, first)
        want = (y0.float() + ref).permute(0, 2, 3, 4, 1)[0]
        assert rel_err(y, want) <= 1e-2
    # AvgDown3D: odd (first chunk: front zero pad) and even frame counts
    for ci, co, ft, fs, t in [(16, 16, 1, 2, 3), (16, 32, 2, 2, 1), (16, 32, 2, 2, 4), (32, 32, 1, 1, 2)]:
        src = torch.randn(1, ci, t, 8, 12, generator=g).bfloat16()
        ref = vae_oracle.avg_down3d(src.float(), co, ft, fs)
        y0 = torch.randn(ref.shape, generator=g).bfloat16()
        y = y0.permute(0, 2, 3, 4, 1)[0].contiguous().cuda()
        ops.avgdown_add_cl(y, src.permute(0, 2, 3, 4, 1)[0].contiguous().cuda(), ft, fs)
        assert rel_err(y, (y0.float() + ref).permute(0, 2, 3, 4, 1)[0]) <= 1e-2
    # patchify / unpatchify + clamp
    x = torch.randn(3, 5, 8, 12, generator=g) * 1.5
    cl = ops.vae_to_cl(x.cuda(), 2, 16)
    ref = vae_oracle._patchify(x[None], 2)[0].permute(1, 2, 3, 0)
    assert torch.equal(cl[..., :12].float().cpu(), ref.bfloat16().float()) and float(cl[..., 12:].float().abs().max()) == 0
    back = torch.zeros(3, 7, 8, 12, device="cuda")
    ops.vae_from_cl(cl.contiguous(), back[:, 1:6], 3, 2, clamp=True)
    assert torch.equal(back[:, 1:6].cpu(), x.bfloat16().float().clamp(-1, 1)) and float(back[:, 0].abs().max()) == 0


def _native(cfg, sd):
    from frameino_b200.vae import AutoencoderKLWan

    m = AutoencoderKLWan(**cfg)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.cuda().eval().prepare()


def _chunks_to_ncthw(frames):
    return torch.cat(frames, dim=0).permute(3, 0, 1, 2)[None].float().cpu()


@pytest.mark.parametrize("lat,h,w", [(3, 4, 6), (1, 4, 6), (4, 8, 4)])
def test_vae_decode_matches_oracle_every_stage(ops, lat, h, w):
    from oracle import vae_oracle

    cfg = synth.VAE_TINY
    sd = synth.make_vae_state_dict(cfg, seed=0)
    z, _ = synth.make_vae_inputs(cfg, lat, h, w, seed=5)
    ref_taps = {}
    with torch.no_grad():
        ref = vae_oracle.decode(sd, cfg, z, taps=ref_taps)
    vae = _native(cfg, sd)
    taps = {}
    vae.__dict__["_fino_taps"] = taps
    out = vae.decode(z.cuda(), return_dict=False)[0]
    assert out.shape == ref.shape and out.dtype == torch.float32
    errs = {k: rel_err(_chunks_to_ncthw(v), ref_taps[k]) for k, v in taps.items()}
    print("vae decode taps:", {k: f"{e:.2e}" for k, e in errs.items()}, "sample", rel_err(out, ref))
    for k, e in errs.items():  # every stage incl. "head" = conv_out before the clamp
        assert e <= TOL, f"{k}: {e}"
    # the sample is clamp(head): with random weights |head| reaches several units, so an error that is <= 2e-2 of the
    # head's range is a larger fraction of the clamped [-1, 1] range; hold it to the head's scale, plus the cosine
    head_max = float(ref_taps["head"].abs().max())
    assert float((out.cpu() - ref).abs().max()) <= TOL * max(1.0, head_max)
    assert cosine(out, ref) >= COS


@pytest.mark.parametrize("lat,h,w", [(3, 4, 6), (1, 4, 6), (2, 8, 4)])
def test_vae_encode_matches_oracle_every_stage(ops, lat, h, w):
    from oracle import vae_oracle

    cfg = synth.VAE_TINY
    sd = synth.make_vae_state_dict(cfg, seed=0)
    _, x = synth.make_vae_inputs(cfg, lat, h, w, seed=5)
    ref_taps = {}
    with torch.no_grad():
        ref = vae_oracle.encode(sd, cfg, x, taps=ref_taps)
    vae = _native(cfg, sd)
    taps = {}
    vae.__dict__["_fino_taps"] = taps
    post = vae.encode(x.cuda()).latent_dist
    assert post.parameters.shape == ref.shape
    errs = {k: rel_err(_chunks_to_ncthw(v), ref_taps[k]) for k, v in taps.items()}
    errs["parameters"] = rel_err(post.parameters, ref)
    print("vae encode taps:", {k: f"{e:.2e}" for k, e in errs.items()})
    for k, e in errs.items():
        assert e <= TOL, f"{k}: {e}"
    assert cosine(post.mode(), ref[:, : cfg["z_dim"]]) >= COS


def test_vae_matches_reference_golden(ops, golden_dir):
    """Against what the reference's own AutoencoderKLWan produced (tests/golden/make_golden.py vae)."""
    g = torch.load(os.path.join(golden_dir, "vae_golden.pt"))
    cfg = synth.VAE_TINY
    vae = _native(cfg, synth.make_vae_state_dict(cfg, seed=0))
    z, x = synth.make_vae_inputs(cfg, 3, 4, 6, seed=5)
    dec = vae.decode(z.cuda(), return_dict=False)[0]
    assert cosine(dec, g["decode.sample"]) >= COS and rel_err(dec, g["decode.sample"]) <= 0.1  # clamped: see above
    dec1 = vae.decode(z[:, :, :1].cuda().bfloat16()).sample  # bf16 latents in -> bf16 video out
    assert dec1.dtype == torch.bfloat16 and cosine(dec1, g["decode1.sample"]) >= COS
    enc = vae.encode(x.cuda(), return_dict=False)[0]
    assert cosine(enc.parameters, g["encode.parameters"]) >= COS and rel_err(enc.parameters, g["encode.parameters"]) <= TOL


def test_vae_wider_channels_and_batch(ops):
    """Channel counts of the real decoder's last two stages (512 -> 256, N tile 256, several K chunks) on a small canvas,
    batch 2."""
    from oracle import vae_oracle

    cfg = dict(synth.VAE_TINY)
    cfg.update(base_dim=64, decoder_base_dim=128)
    sd = synth.make_vae_state_dict(cfg, seed=3)
    z, x = synth.make_vae_inputs(cfg, 2, 2, 4, seed=9)
    z = torch.cat([z, z.flip(3)], dim=0)
    with torch.no_grad():
        ref = vae_oracle.decode(sd, cfg, z)
        ref_e = vae_oracle.encode(sd, cfg, x)
    vae = _native(cfg, sd)
    out = vae.decode(z.cuda(), return_dict=False)[0]
    assert rel_err(out, ref) <= 0.1 and cosine(out, ref) >= COS  # clamped output: see the every-stage test
    enc = vae.encode(x.cuda()).latent_dist.parameters
    assert rel_err(enc, ref_e) <= TOL


def test_vae_full_width_decoder_and_encoder(ops):
    """The real Wan2.2-TI2V-5B VAE widths (decoder 1024/1024/512/256, encoder 160/320/640/640, z 48; 704.7 M parameters)
    on a small canvas (latent 2 x 2 x 4 -> 5 frames of 32 x 64), every stage against the fp32 CPU oracle."""
    from oracle import vae_oracle

    cfg = synth.WAN22_VAE
    sd = synth.make_vae_state_dict(cfg, seed=1)
    z, x = synth.make_vae_inputs(cfg, 2, 2, 4, seed=3)
    ref_taps, ref_taps_e = {}, {}
    with torch.no_grad():
        ref = vae_oracle.decode(sd, cfg, z, taps=ref_taps)
        ref_e = vae_oracle.encode(sd, cfg, x, taps=ref_taps_e)
    vae = _native(cfg, sd)
    taps = {}
    vae.__dict__["_fino_taps"] = taps
    out = vae.decode(z.cuda(), return_dict=False)[0]
    errs = {k: rel_err(_chunks_to_ncthw(v), ref_taps[k]) for k, v in taps.items()}
    print("vae full-width decode taps:", {k: f"{e:.2e}" for k, e in errs.items()})
    assert all(e <= TOL for e in errs.values()), errs
    assert cosine(out, ref) >= COS
    taps.clear()
    enc = vae.encode(x.cuda()).latent_dist.parameters
    errs = {k: rel_err(_chunks_to_ncthw(v), ref_taps_e[k]) for k, v in taps.items()}
    errs["parameters"] = rel_err(enc, ref_e)
    print("vae full-width encode taps:", {k: f"{e:.2e}" for k, e in errs.items()})
    assert all(e <= TOL for e in errs.values()), errs
