"""Denoise-loop glue around the transformer forward, mirroring the reference Wan FrameINO sampler's hot loop
(pipelines/pipeline_wan_i2v_motion_FrameINO.py:809-908): first-frame mask blend, per-token timesteps, frame-wise ID
concat, channel-wise trajectory concat, two CFG forwards, ID-frame drop, flow-match Euler step.

This is the CALLER of the hot path (SURVEY.md §8f "next" #1), in two forms:

* ``wan_frameino_denoise`` — plain torch, drives the native model or any other
  ``transformer(hidden_states=, timestep=, encoder_hidden_states=, return_dict=False)`` callable exactly as the reference
  pipeline does (about 15 small tensor ops per step around the two forwards); parity tests run it on both sides.
* ``wan_frameino_denoise_fused`` — the same loop for the native ``WanTransformer3DModel`` with the glue on the device in
  two kernels per scheduler step (``fino_wan_pack_model_input`` before, ``fino_wan_cfg_euler_step`` after the two
  forwards) and everything step-invariant hoisted out of the loop: the text-embedder MLP and the 30 cross-attention
  K/V projections of both prompts (``WanTextState``), the RoPE tables, the per-token timestep selector. The 5-D model
  input and the 5-D model outputs are never materialised (the forwards run rows -> rows through ``forward_rows``), and
  there is no host synchronisation inside the loop. Same arithmetic, bit for bit, as the plain form.

The scheduler is flow-match Euler with shift 5.0 (config/train_wan_motion_FrameINO.yaml:43-50) — the stepper is not
the thing under test, and it is NOT claimed bit-comparable with a diffusers scheduler object: the scheduler classes are
upstream-only (the checkpoint ships UniPC; the pipeline type-hints FlowMatchEulerDiscreteScheduler), and upstream's
Euler step — recalled — multiplies in the model-output dtype and casts the new sample back to it every step, whereas
this loop keeps the latents and the step in fp32 and applies the static shift once. Everything the reference pipeline
FILE does around the scheduler call is reproduced with its cast points, including the guidance combine in bf16 (:882).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def flow_match_sigmas(num_steps: int, shift: float = 5.0, device="cpu") -> torch.Tensor:
    """sigma_0..sigma_{S-1}, 0  (FlowMatchEulerDiscreteScheduler.set_timesteps with a static shift)."""
    s = torch.linspace(1.0, 1.0 / 1000.0, num_steps, dtype=torch.float32, device=device)
    s = shift * s / (1.0 + (shift - 1.0) * s)
    return torch.cat([s, s.new_zeros(1)])


@torch.no_grad()
def wan_frameino_denoise(
    transformer: Callable,
    latents: torch.Tensor,          # [1, C, F, H, W] initial noise (fp32)
    condition: torch.Tensor,        # [1, C, F, H, W] first-frame latent condition (zeros elsewhere)
    first_frame_mask: torch.Tensor, # [1, C, F, H, W] 0 on latent frame 0, 1 elsewhere (:529-532)
    traj_latents: torch.Tensor,     # [1, C, F + n_id, H, W] trajectory latents, zeros on the ID frames (:516-517)
    id_latents: torch.Tensor,       # [1, C, n_id, H, W]
    prompt_embeds: torch.Tensor,
    negative_prompt_embeds: Optional[torch.Tensor],
    num_steps: int = 50,
    guidance_scale: float = 5.0,
    shift: float = 5.0,
    model_dtype: torch.dtype = torch.bfloat16,
    patch_hw: int = 2,
    cfg_parallel=None,              # frameino_b200.ulysses.CfgParallel: this rank runs one of the two CFG forwards
    callback: Optional[Callable] = None,  # callback(i, t, latents) -> None | replacement latents (pipeline :893-901)
) -> torch.Tensor:
    dev = latents.device
    sigmas = flow_match_sigmas(num_steps, shift, dev)
    n_gen = latents.shape[2]
    n_id = id_latents.shape[2]
    tokens_per_frame = (latents.shape[3] // patch_hw) * (latents.shape[4] // patch_hw)
    do_cfg = guidance_scale > 1.0 and negative_prompt_embeds is not None
    for i in range(num_steps):
        t = sigmas[i] * 1000.0
        x_in = (1 - first_frame_mask) * condition + first_frame_mask * latents  # :829
        ts = (first_frame_mask[0, 0][:, ::patch_hw, ::patch_hw] * t).flatten()  # :842, 0 on frame-0 tokens
        ts = torch.cat([ts, ts.new_full((n_id * tokens_per_frame,), float(t))])[None]  # ID tokens carry t (:834-843)
        x_in = torch.cat([x_in, id_latents.to(x_in.dtype)], dim=2)  # :854 frame-wise
        x_in = torch.cat([x_in, traj_latents.to(x_in.dtype)], dim=1).to(model_dtype)  # :858 channel-wise
        if cfg_parallel is not None:
            if not do_cfg:
                raise ValueError("cfg_parallel needs classifier-free guidance (guidance_scale > 1 and a negative prompt)")
            mine = prompt_embeds if cfg_parallel.branch == 0 else negative_prompt_embeds
            v, vu = cfg_parallel.exchange(
                transformer(hidden_states=x_in, timestep=ts, encoder_hidden_states=mine, return_dict=False)[0])
            v = vu + guidance_scale * (v - vu)  # :882, in the transformer dtype like the reference
        else:
            v = transformer(hidden_states=x_in, timestep=ts, encoder_hidden_states=prompt_embeds, return_dict=False)[0]
        if do_cfg and cfg_parallel is None:
            vu = transformer(hidden_states=x_in, timestep=ts, encoder_hidden_states=negative_prompt_embeds,
                             return_dict=False)[0]
            v = vu + guidance_scale * (v - vu)  # :882, in the transformer dtype (bf16 tensor ops) like the reference
        v = v[:, :, :n_gen].float()  # :886 drop the ID frames
        latents = latents + (sigmas[i + 1] - sigmas[i]) * v  # :891 Euler
        if callback is not None:
            new = callback(i, t, latents)
            latents = latents if new is None else new.to(latents)
    return latents


@torch.no_grad()
def wan_frameino_denoise_fused(
    transformer,                    # frameino_b200.wan.WanTransformer3DModel on a CUDA device
    latents: torch.Tensor,          # [B, C, F, H, W] initial noise
    condition: torch.Tensor,        # [B, C, F, H, W]
    first_frame_mask: torch.Tensor, # [1, 1|C, F, H, W], 0/1 valued (pipeline :529-532)
    traj_latents: torch.Tensor,     # [B, C, F + n_id, H, W]
    id_latents: Optional[torch.Tensor],  # [B, C, n_id, H, W] or None
    prompt_embeds: torch.Tensor,
    negative_prompt_embeds: Optional[torch.Tensor],
    num_steps: int = 50,
    guidance_scale: float = 5.0,
    shift: float = 5.0,
    cfg_parallel=None,              # frameino_b200.ulysses.CfgParallel: this rank runs one of the two CFG forwards
    callback: Optional[Callable] = None,  # callback(i, t, latents) -> None | replacement latents (pipeline :893-901)
) -> torch.Tensor:
    """Same contract and result as ``wan_frameino_denoise`` (see the module docstring for what is fused)."""
    from . import ops

    dev = latents.device
    if dev.type != "cuda":
        raise RuntimeError("frameino_b200 has no CPU path: move the model and inputs to a CUDA device")
    cfg = transformer.config
    patch = tuple(cfg.patch_size)
    p_t, p_h, p_w = patch
    b, c, f, h, w = latents.shape
    n_id = 0 if id_latents is None else id_latents.shape[2]
    if 2 * c != cfg.in_channels or c != cfg.out_channels:
        raise ValueError(f"latent channels {c} do not match the model (in {cfg.in_channels}, out {cfg.out_channels})")
    if p_t != 1:
        raise NotImplementedError("fused sampler loop: temporal patch size 1 (Wan) only")

    f32 = dict(device=dev, dtype=torch.float32)
    lat = latents.to(**f32).clone(memory_format=torch.contiguous_format)
    cond = condition.to(**f32).expand(b, c, f, h, w).contiguous()
    traj = traj_latents.to(**f32).expand(b, c, f + n_id, h, w).contiguous()
    idl = None if n_id == 0 else id_latents.to(**f32).expand(b, c, n_id, h, w).contiguous()
    mask5 = first_frame_mask.to(**f32)
    if mask5.dim() != 5 or mask5.shape[0] != 1 or tuple(mask5.shape[2:]) != (f, h, w):
        raise ValueError(f"first_frame_mask shape {tuple(first_frame_mask.shape)}: want [1, 1|C, {f}, {h}, {w}]")
    mask = mask5[0, 0].contiguous()
    if mask5.shape[1] != 1 and not bool((mask5 == mask5[:, :1]).all()):
        raise NotImplementedError("fused sampler loop: first_frame_mask must be the same for every channel")
    if not bool(((mask == 0) | (mask == 1)).all()):  # one host sync, before the loop
        raise NotImplementedError("fused sampler loop: first_frame_mask must be 0/1 valued (it selects timestep rows)")

    # per-token timestep selector, constant over the loop (pipeline :842-843): mask sampled at the patch corners,
    # ID tokens carry t; time row 0 = timestep 0 (clean first frame), row 1 = t
    sel = mask[:, ::p_h, ::p_w].reshape(-1)
    sel = torch.cat([sel, sel.new_ones(n_id * (h // p_h) * (w // p_w))])
    row_index = sel.to(torch.int32).repeat(b).contiguous()
    mixed = bool((sel == 0).any())  # False: every token carries t (a single time row)

    sigmas = flow_match_sigmas(num_steps, shift, dev)
    sig_host = sigmas.cpu()
    do_cfg = guidance_scale > 1.0 and negative_prompt_embeds is not None
    if cfg_parallel is not None and not do_cfg:
        raise ValueError("cfg_parallel needs classifier-free guidance (guidance_scale > 1 and a negative prompt)")
    run_c = cfg_parallel is None or cfg_parallel.branch == 0
    run_u = do_cfg and (cfg_parallel is None or cfg_parallel.branch == 1)
    text_c = transformer.prepare_text(prompt_embeds) if run_c else None
    text_u = transformer.prepare_text(negative_prompt_embeds) if run_u else None
    grid = (f + n_id, h, w)
    tokens = ((f + n_id) // p_t) * (h // p_h) * (w // p_w)
    rows = None
    for i in range(num_steps):
        t = sigmas[i] * 1000.0
        uniq = torch.stack([torch.zeros_like(t), t]) if mixed else t.reshape(1)
        temb, proj = transformer.time_rows(uniq)
        conditioning = (temb, proj, row_index, 0) if mixed else (temb, proj, None, b * tokens)
        rows = ops.wan_pack_model_input(lat, cond, mask, idl, traj, patch, out=rows)
        y_c = transformer.forward_rows(rows, b, grid, conditioning, text_c) if run_c else None
        y_u = transformer.forward_rows(rows, b, grid, conditioning, text_u) if run_u else None
        if cfg_parallel is not None:  # swap the branch outputs with the partner rank of the other half
            y_c, y_u = cfg_parallel.exchange(y_c if run_c else y_u)
        dsigma = float(sig_host[i + 1] - sig_host[i])  # fp32 difference, as the tensor form computes it
        ops.wan_cfg_euler_step(lat, y_c, y_u, n_id, patch, guidance_scale, dsigma)
        if callback is not None:
            new = callback(i, t, lat)
            if new is not None and new is not lat:
                lat.copy_(new)
    return lat


# ---------------------------------------------------------------------------------------------------------------------
# CogVideoX FrameINO sampler loop glue
# ---------------------------------------------------------------------------------------------------------------------
def dynamic_cfg_scale(guidance_scale: float, num_inference_steps: int, t: float) -> float:
    """pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:906-909 (``use_dynamic_cfg``)."""
    import math

    return 1 + guidance_scale * ((1 - math.cos(math.pi * ((num_inference_steps - t) / num_inference_steps) ** 5.0)) / 2)


def ddim_v_step(model_output: torch.Tensor, sample: torch.Tensor, alpha_t: float, alpha_prev: float) -> torch.Tensor:
    """One deterministic (eta = 0) DDIM step for a v-prediction model — the scheduler family the CogVideoX pipeline runs
    (:915-926). The scheduler classes live upstream in diffusers, not in the reference tree, so the step is the textbook
    one, parameterised by the two cumulative alphas:
        x0 = sqrt(a_t) x - sqrt(1 - a_t) v;   eps = sqrt(a_t) v + sqrt(1 - a_t) x;
        x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps."""
    a_t, a_p = float(alpha_t), float(alpha_prev)
    x0 = a_t ** 0.5 * sample - (1 - a_t) ** 0.5 * model_output
    eps = a_t ** 0.5 * model_output + (1 - a_t) ** 0.5 * sample
    return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


def scaled_linear_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                                 snr_shift_scale: float = 1.0) -> torch.Tensor:
    """``scaled_linear`` betas + the SNR shift of the CogVideoX schedulers (upstream defaults, recalled; pass your own
    table to ``cog_frameino_denoise`` to use another schedule)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float64) ** 2
    ac = torch.cumprod(1.0 - betas, dim=0)
    ac = ac / (snr_shift_scale + (1 - snr_shift_scale) * ac)
    return ac.float()


def rescale_zero_terminal_snr(alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """Zero-terminal-SNR rescale of a cumulative-alpha table (Lin et al. 2023, algorithm 1), as the CogVideoX schedulers
    apply it to ``alphas_cumprod`` (upstream diffusers, recalled)."""
    s = alphas_cumprod.double().sqrt()
    s0, st = s[0].clone(), s[-1].clone()
    s = (s - st) * (s0 / (s0 - st))
    return (s ** 2).to(alphas_cumprod.dtype)


class CogVideoXDPMSchedule:
    """The arithmetic of diffusers' ``CogVideoXDPMScheduler`` that the FrameINO CogVideoX pipeline drives
    (pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:30, :915-926: ``step(noise_pred, old_pred_original_sample, t,
    timesteps[i-1] if i > 0 else None, latents, ...) -> (prev_sample, pred_original_sample)``). The class itself lives
    upstream in diffusers, not in the reference tree, so this is a restatement of its published algorithm from recalled
    semantics: DPM-Solver++(2M) SDE on the VP schedule — ``scaled_linear`` betas 0.00085..0.012, SNR shift
    (``snr_shift_scale`` 1.0 for CogVideoX-5B, 3.0 for 2B), zero-terminal-SNR rescale, ``trailing`` timestep spacing,
    v-prediction. ``tests/test_cog_loop.py`` checks it against the closed form of the paper (Lu et al. 2022, eq. for the
    second-order multistep SDE solver) written in (alpha, sigma, lambda) form."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 snr_shift_scale: float = 1.0, rescale_betas_zero_snr: bool = True, prediction_type: str = "v_prediction"):
        ac = scaled_linear_alphas_cumprod(num_train_timesteps, beta_start, beta_end, snr_shift_scale)
        self.alphas_cumprod = rescale_zero_terminal_snr(ac) if rescale_betas_zero_snr else ac
        self.final_alpha_cumprod = torch.tensor(1.0)  # set_alpha_to_one
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.num_inference_steps = None
        self.timesteps = None
        self.order = 1
        self.init_noise_sigma = 1.0

    def set_timesteps(self, num_inference_steps: int, device=None) -> torch.Tensor:
        """``timestep_spacing="trailing"``: round(arange(T, 0, -T/steps)) - 1."""
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps / num_inference_steps
        ts = torch.round(torch.arange(self.num_train_timesteps, 0, -ratio, dtype=torch.float64)).long() - 1
        self.timesteps = ts.to(device) if device is not None else ts
        return self.timesteps

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    @staticmethod
    def _lambda(alpha_prod: torch.Tensor) -> torch.Tensor:
        return ((alpha_prod / (1 - alpha_prod)) ** 0.5).log()

    def coefficients(self, timestep: int, timestep_back: Optional[int]):
        """(alpha_t, alpha_prev, mult1, mult2, mult3 | None, mult4 | None, mult_noise, prev_timestep) as 0-d tensors."""
        prev_timestep = int(timestep) - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[int(timestep)]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        a_back = self.alphas_cumprod[int(timestep_back)] if timestep_back is not None else None
        lamb, lamb_next = self._lambda(a_t), self._lambda(a_prev)
        h = lamb_next - lamb
        mult1 = ((1 - a_prev) / (1 - a_t)) ** 0.5 * (-h).exp()
        mult2 = (-2 * h).expm1() * a_prev ** 0.5
        mult3 = mult4 = None
        if a_back is not None:
            r = (lamb - self._lambda(a_back)) / h
            mult3, mult4 = 1 + 1 / (2 * r), 1 / (2 * r)
        mult_noise = (1 - a_prev) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        return a_t, a_prev, mult1, mult2, mult3, mult4, mult_noise, prev_timestep

    def step(self, model_output: torch.Tensor, old_pred_original_sample: Optional[torch.Tensor], timestep,
             timestep_back, sample: torch.Tensor, eta: float = 0.0, generator=None, noise: Optional[torch.Tensor] = None,
             return_dict: bool = False):
        """One solver step. ``noise``: the standard-normal draw to use (else drawn from ``generator`` on the sample's
        device); the SDE solver adds ``mult_noise * noise`` every step (it vanishes on the last one)."""
        if self.num_inference_steps is None:
            raise ValueError("set_timesteps must run before step")
        a_t, a_prev, m1, m2, m3, m4, m_noise, prev_timestep = self.coefficients(int(timestep), timestep_back)
        if self.prediction_type == "v_prediction":
            x0 = a_t ** 0.5 * sample - (1 - a_t) ** 0.5 * model_output
        elif self.prediction_type == "epsilon":
            x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            raise ValueError(f"prediction_type {self.prediction_type!r}")
        if noise is None:
            noise = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=sample.dtype)
        if old_pred_original_sample is None or prev_timestep < 0:
            prev = m1 * sample - m2 * x0 + m_noise * noise  # first step and last step: first order
        else:
            denoised = m3 * x0 - m4 * old_pred_original_sample
            prev = m1 * sample - m2 * denoised + m_noise * noise
        return prev, x0


@torch.no_grad()
def cog_frameino_denoise(
    transformer: Callable,
    latents: torch.Tensor,          # [1, F, C, H, W] initial noise (CogVideoX layout: frames before channels)
    image_latents: torch.Tensor,    # [1, F, C, H, W] first-frame condition (zeros after frame 0)
    traj_latents: torch.Tensor,     # [1, F, C, H, W] trajectory latents
    id_latent: Optional[torch.Tensor],  # [1, n_id, C, H, W] or None
    prompt_embeds: torch.Tensor,    # [2, T, text_dim] = cat(negative, positive) when guidance is on (:767-768), else [1, ...]
    image_rotary_emb,
    timesteps,                      # decreasing integer timesteps, e.g. torch.linspace(999, 0, steps).long()
    alphas_cumprod: Optional[torch.Tensor] = None,   # [num_train_timesteps] (DDIM form); unused with ``scheduler``
    guidance_scale: float = 6.0,
    use_dynamic_cfg: bool = False,
    model_dtype: torch.dtype = torch.bfloat16,
    scheduler: Optional[CogVideoXDPMSchedule] = None,  # the pipeline's CogVideoXDPMScheduler branch (:918-926)
    generator=None,
) -> torch.Tensor:
    """The hot loop of pipelines/pipeline_cogvideox_i2v_motion_FrameINO.py:846-927 around the transformer forward:
    batched CFG (:853, :856, :859), frame-wise ID concat with zero padding of the image / trajectory streams (:862-873),
    channel-wise concat (:877), ID-frame drop (:899-900), dynamic guidance (:904-909), CFG combine (:910-912), scheduler
    step — ``CogVideoXDPMSchedule.step`` with its two history arguments when ``scheduler`` is given (:918-926), else the
    DDIM-style ``ddim_v_step`` (:916) — and the cast back to the prompt dtype (:927). Plain torch: drives the native model
    or any callable with the reference forward signature."""
    do_cfg = guidance_scale > 1.0
    steps = len(timesteps)
    n_frames = latents.shape[1]
    lat = latents
    old_x0 = None  # "for DPM-solver++" (:848)
    for i in range(steps):
        t = timesteps[i]
        x = torch.cat([lat] * 2) if do_cfg else lat
        img = torch.cat([image_latents] * 2) if do_cfg else image_latents
        traj = torch.cat([traj_latents] * 2) if do_cfg else traj_latents
        if id_latent is not None:
            idl = torch.cat([id_latent] * 2) if do_cfg else id_latent
            x = torch.cat([x, idl.to(x.dtype)], dim=1)
            pad = x.new_zeros(idl.shape)
            img = torch.cat([img.to(x.dtype), pad], dim=1)
            traj = torch.cat([traj.to(x.dtype), pad], dim=1)
        x = torch.cat([x, img.to(x.dtype), traj.to(x.dtype)], dim=2)
        ts = torch.as_tensor(t, device=x.device).expand(x.shape[0])
        v = transformer(hidden_states=x.to(model_dtype), encoder_hidden_states=prompt_embeds, timestep=ts,
                        image_rotary_emb=image_rotary_emb, return_dict=False)[0].float()
        if id_latent is not None:
            v = v[:, :n_frames]
        g = dynamic_cfg_scale(guidance_scale, steps, float(t)) if use_dynamic_cfg else guidance_scale
        if do_cfg:
            v_uncond, v_text = v.chunk(2)
            v = v_uncond + g * (v_text - v_uncond)
        if scheduler is not None:
            lat_new, old_x0 = scheduler.step(v, old_x0, int(t), int(timesteps[i - 1]) if i > 0 else None, lat.float(),
                                             generator=generator)
            lat = lat_new.to(prompt_embeds.dtype)
        else:
            a_t = alphas_cumprod[int(t)]
            a_prev = alphas_cumprod[int(timesteps[i + 1])] if i + 1 < steps else torch.tensor(1.0)
            lat = ddim_v_step(v, lat.float(), a_t, a_prev).to(prompt_embeds.dtype)
    return lat
