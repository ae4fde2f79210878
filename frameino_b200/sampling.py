"""Denoise-loop glue around the transformer forward, mirroring the reference Wan FrameINO sampler's hot loop
(pipelines/pipeline_wan_i2v_motion_FrameINO.py:809-908): first-frame mask blend, per-token timesteps, frame-wise ID
concat, channel-wise trajectory concat, two CFG forwards, ID-frame drop, flow-match Euler step.

This is the CALLER of the hot path (SURVEY.md §8f "next" #1), kept in plain torch so that the same function can drive
either the native model or any other ``transformer(hidden_states=, timestep=, encoder_hidden_states=, return_dict=False)``
callable; parity tests run it on both sides. The scheduler is flow-match Euler with shift 5.0
(config/train_wan_motion_FrameINO.yaml:43-50) — the stepper is not the thing under test.
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
        v = transformer(hidden_states=x_in, timestep=ts, encoder_hidden_states=prompt_embeds, return_dict=False)[0]
        if do_cfg:
            vu = transformer(hidden_states=x_in, timestep=ts, encoder_hidden_states=negative_prompt_embeds,
                             return_dict=False)[0]
            v = vu.float() + guidance_scale * (v.float() - vu.float())  # :882
        v = v[:, :, :n_gen].float()  # :886 drop the ID frames
        latents = latents + (sigmas[i + 1] - sigmas[i]) * v  # :891 Euler
    return latents
