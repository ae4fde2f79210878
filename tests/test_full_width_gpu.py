"""-m gpu: parity AT THE BENCHMARKED WIDTH. One full-width Wan2.2-5B block (D 3072, FFN 14336, 24 heads x 128,
per-token modulation) over 3520 tokens and one full-width CogVideoX-5B block (48 heads x 64, joint text + video,
S = 226 + 4050) run through the native models with ``num_layers=1`` — i.e. patch embed, conditioning, the block, the
output head — against the CPU oracle evaluated on the same bf16 weights, with EVERY tapped tensor inside the block held
to the north-star bar (max|a-b| / max|b| <= 2e-2). The K = 14336 FFN-down accumulation with its gate*residual epilogue,
the 3072-wide RMSNorm-across-heads and the 24 x 128 / 48 x 64 attention shapes are the ones bench.py times.
(transformer_wan.py:308-350, cogvideox_transformer_3d.py:122-161.)"""
import pytest
import torch

from conftest import cosine, rel_err
from frameino_b200 import synth

pytestmark = pytest.mark.gpu
LAYER_TOL = 2e-2
COS_TOL = 0.999


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.mark.parametrize("per_token", [True, False])
def test_wan_full_width_block_every_tap(per_token):
    from frameino_b200.wan import WanTransformer3DModel
    from oracle import wan_oracle

    cfg = dict(synth.WAN22_5B)
    cfg["num_layers"] = 1
    sd = synth.make_wan_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    # latent 3 + 1 ID frames x 44 x 80 -> 4 * 22 * 40 = 3520 tokens (one 8-GPU shard of config 2), 512 text tokens
    hidden, ts, text = synth.make_wan_inputs(cfg, 3, 44, 80, n_id=1, text_len=512, text_true_len=120,
                                             per_token_timestep=per_token, dtype=torch.bfloat16)
    ref_taps = {}
    ref = wan_oracle.wan_forward(sd, wan_oracle.WanConfig(**cfg), hidden, ts, text, taps=ref_taps)
    model = WanTransformer3DModel(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to_inference_dtype(torch.bfloat16).cuda().eval()
    taps = {}
    model.__dict__["_fino_taps"] = taps
    out = model(hidden_states=hidden.cuda(), timestep=ts.cuda(), encoder_hidden_states=text.cuda(), return_dict=False)[0]
    assert torch.isfinite(out.float()).all()
    errs = {}
    for name in ("patch_embed", "text", "blocks.0.norm1", "blocks.0.after_attn1", "blocks.0.after_attn2", "blocks.0.out"):
        errs[name] = rel_err(taps[name], ref_taps[name])
    errs["sample"] = rel_err(out, ref)
    print("wan full-width taps:", {k: f"{v:.2e}" for k, v in errs.items()})
    for name, e in errs.items():
        assert e <= LAYER_TOL, f"{name}: {e}"
    assert cosine(out, ref) >= COS_TOL
    assert cosine(taps["blocks.0.out"], ref_taps["blocks.0.out"]) >= COS_TOL


@pytest.mark.parametrize("batch", [1, 2])
def test_cog_full_width_block_every_tap(batch):
    from frameino_b200.cogvideox import CogVideoXTransformer3DModel
    from oracle import cog_oracle

    cfg = dict(synth.COG_5B_I2V)
    cfg["num_layers"] = 1
    cfg["sample_frames"] = 5  # 2 latent frames + 1 ID frame -> 3 * 30 * 45 = 4050 video tokens + 226 text = 4276
    lat_f = 2
    h, w = cfg["sample_height"], cfg["sample_width"]
    sd = synth.make_cog_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    hidden, ts, text = synth.make_cog_inputs(cfg, lat_f, h, w, n_id=1, batch=batch, dtype=torch.bfloat16)
    cos, sin = cog_oracle.cog_rope_3d(64, (h // 2, w // 2), lat_f, 1)
    ref_taps = {}
    ref = cog_oracle.cog_forward(sd, cfg, hidden, text, ts, (cos, sin), taps=ref_taps)
    model = CogVideoXTransformer3DModel(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to_inference_dtype(torch.bfloat16).cuda().eval()
    taps = {}
    model.__dict__["_fino_taps"] = taps
    out = model(hidden_states=hidden.cuda(), encoder_hidden_states=text.cuda(), timestep=ts.cuda(),
                image_rotary_emb=(cos.cuda(), sin.cuda()), return_dict=False)[0]
    assert torch.isfinite(out.float()).all()
    p = "transformer_blocks.0"
    errs = {n: rel_err(taps[f"{p}.{n}"], ref_taps[f"{p}.{n}"]) for n in ("after_attn", "out", "enc")}
    errs["sample"] = rel_err(out, ref)
    print("cog full-width taps:", {k: f"{v:.2e}" for k, v in errs.items()})
    for name, e in errs.items():
        assert e <= LAYER_TOL, f"{name}: {e}"
    assert cosine(out, ref) >= COS_TOL
