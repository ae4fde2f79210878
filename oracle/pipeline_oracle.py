"""CPU oracle (test infrastructure) — the Wan FrameINO image-to-video pipeline around the two hot paths, restated from
``/root/reference/pipelines/pipeline_wan_i2v_motion_FrameINO.py`` (``prepare_latents`` :400-553 and ``__call__`` :766-945,
the Wan2.2 ``expand_timesteps`` branch) as pure functions over the Wan and VAE state dicts.

Only ``tests/`` may import this file. What is NOT restated because it is not under ``/root/reference``:

  * the text encoder (UMT5) — ``prompt_embeds`` / ``negative_prompt_embeds`` are inputs, as the pipeline accepts (:687-688);
  * the scheduler object — flow-match Euler with the static shift of ``config/train_wan_motion_FrameINO.yaml:43-50`` is
    used on both sides of every comparison (see ``frameino_b200/sampling.py`` for why);
  * ``VideoProcessor`` (upstream diffusers, recalled): ``preprocess`` of a tensor in [0, 1] is ``2 x - 1``,
    ``postprocess_video`` is ``(x / 2 + 0.5).clamp(0, 1)`` per frame, ``"pt"`` -> [B, F, C, H, W], ``"np"`` -> [B, F, H, W, C].

Pinned: tests/test_oracle_golden.py compares ``prepare_latents`` and ``generate`` with outputs of the reference's OWN
pipeline file executed on CPU around the reference transformer and VAE (tests/golden/pipeline_golden.pt, written by
``make_golden.py pipeline`` through the diffusers shim): 1e-5 / 1e-4. The three upstream-only pieces above are the shim's
restatements on the reference side of that comparison too, so for THEM parity stays "unpinned"; every line taken from
the pipeline file itself is restated with its cast points, cited, and pinned.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import vae_oracle, wan_oracle


def _stats(vae_cfg: dict, like: torch.Tensor):
    """pipeline :448-455 — note ``latents_std`` is the RECIPROCAL of the config's list."""
    z = vae_cfg["z_dim"]
    mean = torch.tensor(vae_cfg["latents_mean"]).view(1, z, 1, 1, 1).to(like.device, like.dtype)
    inv_std = 1.0 / torch.tensor(vae_cfg["latents_std"]).view(1, z, 1, 1, 1).to(like.device, like.dtype)
    return mean, inv_std


def _encode_mode(vae_sd, vae_cfg, x: torch.Tensor) -> torch.Tensor:
    """``retrieve_latents(vae.encode(x), sample_mode="argmax")`` (:464): the posterior mean."""
    return vae_oracle.encode(vae_sd, vae_cfg, x)[:, : vae_cfg["z_dim"]]


def prepare_latents(vae_sd: Dict[str, torch.Tensor], vae_cfg: dict, image: torch.Tensor, traj_tensor: torch.Tensor,
                    id_tensor: Optional[torch.Tensor], batch_size: int, height: int, width: int, num_frames: int,
                    latents: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                    dtype: torch.dtype = torch.float32):
    """pipeline :400-536. image [B', 3, H, W] in [-1, 1]; traj_tensor [F, 3, H, W]; id_tensor [B', 3, n_id, H, W] or
    None -> (latents, latent_condition, traj_latents, id_latent_condition | None, first_frame_mask)."""
    s_t, s_s = vae_cfg["scale_factor_temporal"], vae_cfg["scale_factor_spatial"]
    f_lat = (num_frames - 1) // s_t + 1  # :416
    h_lat, w_lat = height // s_s, width // s_s
    shape = (batch_size, vae_cfg["z_dim"], f_lat, h_lat, w_lat)
    if latents is None:
        latents = torch.randn(shape, generator=generator, dtype=dtype)  # :427-428
    else:
        latents = latents.to(dtype)
    video_condition = image.unsqueeze(2).float()  # :432-435 (expand_timesteps: the single first frame)
    mean, inv_std = _stats(vae_cfg, latents)
    cond = _encode_mode(vae_sd, vae_cfg, video_condition).repeat(batch_size, 1, 1, 1, 1)  # :464-465
    cond = (cond.to(dtype) - mean) * inv_std  # :467-468
    traj = traj_tensor.float().unsqueeze(0).permute(0, 2, 1, 3, 4)  # :473-475
    traj_lat = (_encode_mode(vae_sd, vae_cfg, traj) - mean) * inv_std  # :478-481
    traj_lat = traj_lat.contiguous().float()  # :484
    id_cond = None
    if id_tensor is not None and id_tensor.shape[2] != 0:  # :489
        parts = []
        for k in range(id_tensor.shape[2]):  # :497-511, one single-frame encode per ID frame
            lat = _encode_mode(vae_sd, vae_cfg, id_tensor[:, :, k].unsqueeze(2).float()).repeat(batch_size, 1, 1, 1, 1)
            parts.append((lat.to(dtype) - mean) * inv_std)
        id_cond = torch.cat(parts, dim=2)  # :514
        traj_lat = torch.cat([traj_lat, torch.zeros_like(id_cond)], dim=2)  # :517-518
    mask = torch.ones(1, 1, f_lat, h_lat, w_lat, dtype=dtype)  # :529-532
    mask[:, :, 0] = 0
    return latents, cond, traj_lat, id_cond, mask


def flow_match_sigmas(num_steps: int, shift: float) -> torch.Tensor:
    s = torch.linspace(1.0, 1.0 / 1000.0, num_steps, dtype=torch.float32)
    s = shift * s / (1.0 + (shift - 1.0) * s)
    return torch.cat([s, s.new_zeros(1)])


@torch.no_grad()
def generate(wan_sd, wan_cfg: "wan_oracle.WanConfig", vae_sd, vae_cfg: dict, image: torch.Tensor,
             traj_tensor: torch.Tensor, id_tensor: Optional[torch.Tensor], prompt_embeds: torch.Tensor,
             negative_prompt_embeds: Optional[torch.Tensor], height: int, width: int, num_frames: int,
             num_inference_steps: int = 50, guidance_scale: float = 5.0, shift: float = 5.0,
             latents: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
             output_type: str = "pt", transformer_dtype: torch.dtype = torch.float32, taps: Optional[dict] = None):
    """pipeline ``__call__`` :766-945 (image already pre-processed to [-1, 1], prompts already embedded)."""
    batch = prompt_embeds.shape[0]  # :729
    prompt_embeds = prompt_embeds.to(transformer_dtype)  # :746-749
    if negative_prompt_embeds is not None:
        negative_prompt_embeds = negative_prompt_embeds.to(transformer_dtype)
    lat, cond, traj, id_cond, mask = prepare_latents(vae_sd, vae_cfg, image, traj_tensor, id_tensor, batch, height,
                                                     width, num_frames, latents, generator)  # :773-787
    n_gen, h_lat, w_lat = lat.shape[2], lat.shape[3], lat.shape[4]  # :796
    do_cfg = guidance_scale > 1.0  # :555-557
    sigmas = flow_match_sigmas(num_inference_steps, shift)
    for i in range(num_inference_steps):
        t = sigmas[i] * 1000.0
        x = ((1 - mask) * cond + mask * lat).to(transformer_dtype)  # :829-830
        mask_adj = mask if id_cond is None else torch.cat(
            [mask, torch.ones(1, 1, id_cond.shape[2], h_lat, w_lat, dtype=transformer_dtype)], dim=2)  # :833-839
        timestep = (mask_adj[0][0][:, ::2, ::2] * t).flatten().unsqueeze(0).expand(lat.shape[0], -1)  # :842-843
        if id_cond is not None:
            x = torch.cat([x, id_cond.to(x.dtype)], dim=2)  # :853-854
        x = torch.cat([x, traj.to(x.dtype)], dim=1).to(transformer_dtype)  # :858
        v = wan_oracle.wan_forward(wan_sd, wan_cfg, x, timestep.float(), prompt_embeds)  # :863-870
        if do_cfg:
            vu = wan_oracle.wan_forward(wan_sd, wan_cfg, x, timestep.float(), negative_prompt_embeds)  # :874-881
            v = vu + guidance_scale * (v - vu)  # :882
        v = v[:, :, :n_gen]  # :886
        lat = lat + (sigmas[i + 1] - sigmas[i]) * v.float()  # :891 (Euler flow-match, fp32 latents)
    lat = (1 - mask) * cond + mask * lat  # :914-915
    if taps is not None:
        taps["latents"] = lat
    if output_type == "latent":  # :931-932
        return lat
    mean, inv_std = _stats(vae_cfg, lat)
    video = vae_oracle.decode(vae_sd, vae_cfg, lat / inv_std + mean)  # :917-928
    video = (video / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4)  # VideoProcessor.postprocess_video -> [B, F, C, H, W]
    if output_type == "np":
        return video.permute(0, 1, 3, 4, 2).float().numpy()
    return video
