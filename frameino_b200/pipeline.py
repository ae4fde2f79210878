"""The caller on either side of the two hot paths: the Wan FrameINO image-to-video pipeline
(``/root/reference/pipelines/pipeline_wan_i2v_motion_FrameINO.py``: ``prepare_latents`` :400-553, ``__call__`` :766-945),
Wan2.2 ``expand_timesteps`` form, with every stage on the device:

  pixels --3 x AutoencoderKLWan.encode--> latents --wan_frameino_denoise_fused (50 x 2 forwards)--> latents
         --AutoencoderKLWan.decode--> pixels

``WanFrameINOPipeline.__call__`` keeps the reference's keyword names, defaults, validation errors and return type
(``.frames`` / a 1-tuple), so that ``pipe(image=..., traj_tensor=..., ID_tensor=..., prompt_embeds=..., ...)`` reads like
the reference's call (app.py:705-714). Out of scope, and said so loudly instead of being faked:

  * the UMT5 text encoder / tokenizer: ``prompt=`` strings need a ``text_encoder`` callable handed to the constructor
    (``text_encoder(list_of_str, max_sequence_length) -> [B, T, text_dim]``); otherwise pass ``prompt_embeds`` /
    ``negative_prompt_embeds`` as the reference allows (:687-688);
  * the Wan2.1 branches (CLIP image embeds, ``last_image``, two-stage ``transformer_2`` / ``boundary_ratio``): the
    FrameINO checkpoints are Wan2.2-TI2V-5B (``expand_timesteps=True``, app.py:150-156);
  * the scheduler object: flow-match Euler with a static shift (``sampling.py`` explains the deviation).

There is no CPU path: every stage raises when the models are not on a CUDA device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Union

import logging

import torch

from .modules import ConfigDict
from .sampling import wan_frameino_denoise, wan_frameino_denoise_fused


logger = logging.getLogger(__name__)


@dataclass
class WanPipelineOutput:
    """diffusers.pipelines.wan.pipeline_output.WanPipelineOutput"""
    frames: Any


def retrieve_latents(encoder_output, generator: Optional[torch.Generator] = None, sample_mode: str = "sample"):
    """pipeline :111-122"""
    if hasattr(encoder_output, "latent_dist") and sample_mode == "sample":
        return encoder_output.latent_dist.sample(generator)
    if hasattr(encoder_output, "latent_dist") and sample_mode == "argmax":
        return encoder_output.latent_dist.mode()
    if hasattr(encoder_output, "latents"):
        return encoder_output.latents
    raise AttributeError("Could not access latents of provided encoder_output")


class VideoProcessor:
    """The two calls the pipeline makes on diffusers' ``VideoProcessor`` (upstream, recalled): ``preprocess`` (:767) and
    ``postprocess_video`` (:929)."""

    def __init__(self, vae_scale_factor: int = 8):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height: Optional[int] = None, width: Optional[int] = None) -> torch.Tensor:
        """-> [B, 3, H, W] fp32 in [-1, 1]. Tensors / arrays are taken as [0, 1] unless they hold negative values
        (then they are taken as already normalised, as upstream does with a warning); PIL images are resized (Lanczos)."""
        try:
            import PIL.Image
        except ImportError:  # pragma: no cover
            PIL = None
        if PIL is not None and isinstance(image, PIL.Image.Image):
            import numpy as np

            if height is not None and width is not None and image.size != (width, height):
                image = image.resize((width, height), PIL.Image.LANCZOS)
            image = torch.from_numpy(np.asarray(image.convert("RGB"), dtype=np.float32) / 255.0).permute(2, 0, 1)
        elif not isinstance(image, torch.Tensor):
            import numpy as np

            arr = np.asarray(image, dtype=np.float32)
            image = torch.from_numpy(arr).permute(2, 0, 1) if arr.ndim == 3 else torch.from_numpy(arr).permute(0, 3, 1, 2)
        if image.dim() == 3:
            image = image.unsqueeze(0)
        if image.dim() != 4 or image.shape[1] != 3:
            raise ValueError(f"image must be [3, H, W] or [B, 3, H, W], got {tuple(image.shape)}")
        if height is not None and width is not None and tuple(image.shape[-2:]) != (height, width):
            raise ValueError(f"image is {tuple(image.shape[-2:])}, the call asks for {(height, width)}: resize it first")
        image = image.float()
        if image.numel() and float(image.min()) < 0:
            return image
        return 2.0 * image - 1.0

    @staticmethod
    def postprocess_video(video: torch.Tensor, output_type: str = "np"):
        """video [B, C, F, H, W] in [-1, 1] -> ``"pt"`` [B, F, C, H, W] in [0, 1] / ``"np"`` [B, F, H, W, C] /
        ``"pil"`` list of lists of images."""
        v = (video / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4)
        if output_type == "pt":
            return v
        arr = v.permute(0, 1, 3, 4, 2).float().cpu().numpy()
        if output_type == "np":
            return arr
        if output_type == "pil":
            import PIL.Image

            return [[PIL.Image.fromarray((f * 255).round().astype("uint8")) for f in vid] for vid in arr]
        raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil', 'latent']")


class WanFrameINOPipeline:
    """Stand-in for the reference ``WanImageToVideoPipeline`` (FrameINO variant) — see the module docstring."""

    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]

    def __init__(self, vae, transformer, text_encoder: Optional[Callable] = None, scheduler=None,
                 expand_timesteps: bool = True, boundary_ratio: Optional[float] = None, shift: float = 5.0):
        if not expand_timesteps:
            raise NotImplementedError("only the Wan2.2 (expand_timesteps=True) pipeline form is built — the FrameINO "
                                      "checkpoints are Wan2.2-TI2V-5B")
        if boundary_ratio is not None:
            raise NotImplementedError("two-stage denoising (transformer_2 / boundary_ratio) is a Wan2.2-A14B feature")
        if scheduler is not None:  # only its static shift is read: the stepper is flow-match Euler (sampling.py)
            sched_cfg = getattr(scheduler, "config", scheduler)
            shift = sched_cfg["shift"] if isinstance(sched_cfg, dict) else getattr(sched_cfg, "shift", None)
            if shift is None:
                raise NotImplementedError("the sampler is flow-match Euler with a static shift; the scheduler object "
                                          "handed in carries no `shift`")
        self.vae, self.transformer, self.text_encoder, self.scheduler = vae, transformer, text_encoder, scheduler
        self.shift = float(shift)
        self.config = ConfigDict(dict(boundary_ratio=boundary_ratio, expand_timesteps=expand_timesteps))
        self.vae_scale_factor_temporal = vae.config.scale_factor_temporal if vae is not None else 4  # :200
        self.vae_scale_factor_spatial = vae.config.scale_factor_spatial if vae is not None else 8  # :201
        self.video_processor = VideoProcessor(vae_scale_factor=self.vae_scale_factor_spatial)
        self._guidance_scale = 5.0
        self._num_timesteps = 0
        self._current_timestep = None
        self._interrupt = False
        self._attention_kwargs = None

    # ---- the reference's read-only properties (:555-580) -------------------------------------------------------------
    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1.0

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def current_timestep(self):
        return self._current_timestep

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def attention_kwargs(self):
        return self._attention_kwargs

    @property
    def _execution_device(self) -> torch.device:
        return self.transformer.device

    def to(self, *args, **kwargs) -> "WanFrameINOPipeline":
        self.vae.to(*args, **kwargs)
        self.transformer.to(*args, **kwargs)
        return self

    # ---- :339-397 ----------------------------------------------------------------------------------------------------
    def check_inputs(self, prompt, negative_prompt, image, height, width, prompt_embeds=None,
                     negative_prompt_embeds=None, image_embeds=None, callback_on_step_end_tensor_inputs=None,
                     guidance_scale_2=None):
        if image is not None and image_embeds is not None:
            raise ValueError("Cannot forward both `image` and `image_embeds`. Please make sure to only forward one of "
                             "the two.")
        if image is None and image_embeds is None:
            raise ValueError("Provide either `image` or `prompt_embeds`. Cannot leave both `image` and `image_embeds` "
                             "undefined.")
        if height % 16 != 0 or width % 16 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 16 but are {height} and {width}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            bad = [k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but "
                             f"found {bad}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`. Please make sure to only forward one of "
                             "the two.")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `negative_prompt` and `negative_prompt_embeds`. Please make sure to "
                             "only forward one of the two.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and "
                             "`prompt_embeds` undefined.")
        if prompt is not None and not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and not isinstance(negative_prompt, (str, list)):
            raise ValueError(f"`negative_prompt` has to be of type `str` or `list` but is {type(negative_prompt)}")
        if guidance_scale_2 is not None:  # boundary_ratio is always None here
            raise ValueError("`guidance_scale_2` is only supported when the pipeline's `boundary_ratio` is not None.")

    # ---- :236-337 ----------------------------------------------------------------------------------------------------
    def encode_prompt(self, prompt, negative_prompt=None, do_classifier_free_guidance: bool = True,
                      num_videos_per_prompt: int = 1, prompt_embeds=None, negative_prompt_embeds=None,
                      max_sequence_length: int = 512, device=None, dtype=None):
        def embed(p):
            if self.text_encoder is None:
                raise NotImplementedError(
                    "the UMT5 text encoder is outside frameino_b200: pass prompt_embeds / negative_prompt_embeds, or "
                    "construct the pipeline with text_encoder=callable(list[str], max_sequence_length) -> [B, T, D]")
            e = self.text_encoder([p] if isinstance(p, str) else list(p), max_sequence_length)
            b, t, _ = e.shape
            return e.repeat(1, num_videos_per_prompt, 1).view(b * num_videos_per_prompt, t, -1)  # :228-229

        if prompt is not None:  # :280-283
            batch = 1 if isinstance(prompt, str) else len(prompt)
        else:
            batch = prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_embeds = embed(prompt)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            neg = negative_prompt or ""  # :314
            neg = batch * [neg] if isinstance(neg, str) else neg
            if len(neg) != batch:
                raise ValueError(f"`negative_prompt` has batch size {len(neg)}, but `prompt` has batch size {batch}.")
            negative_prompt_embeds = embed(neg)
        kw = {k: v for k, v in dict(device=device, dtype=dtype).items() if v is not None}
        prompt_embeds = prompt_embeds.to(**kw)
        if negative_prompt_embeds is not None:
            negative_prompt_embeds = negative_prompt_embeds.to(**kw)
        return prompt_embeds, negative_prompt_embeds

    # ---- :400-536 ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare_latents(self, image: torch.Tensor, traj_tensor: torch.Tensor, ID_tensor: Optional[torch.Tensor],
                        batch_size: int, num_channels_latents: int = 16, height: int = 480, width: int = 832,
                        num_frames: int = 81, dtype: Optional[torch.dtype] = None,
                        device: Optional[torch.device] = None, generator=None, latents: Optional[torch.Tensor] = None,
                        last_image: Optional[torch.Tensor] = None):
        """-> (latents, latent_condition, traj_latents, ID_latent_condition | None, first_frame_mask), the five values
        of the reference's ``expand_timesteps`` return (:536). Three (or 2 + n_id) VAE encodes, all on the device."""
        if last_image is not None:
            raise NotImplementedError("last_image is a Wan2.1 FLF2V input")
        dtype = dtype or torch.float32
        device = torch.device(device) if device is not None else self._execution_device
        vae, z = self.vae, self.vae.config.z_dim
        f_lat = (num_frames - 1) // self.vae_scale_factor_temporal + 1  # :416
        h_lat, w_lat = height // self.vae_scale_factor_spatial, width // self.vae_scale_factor_spatial
        shape = (batch_size, num_channels_latents, f_lat, h_lat, w_lat)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}. Make sure the batch size matches the length of "
                             f"the generators.")
        if latents is None:  # diffusers randn_tensor: drawn on the generator's device, then moved
            gens = generator if isinstance(generator, list) else [generator] * batch_size
            one = (1,) + shape[1:]
            if isinstance(generator, list):
                latents = torch.cat([torch.randn(one, generator=g, device=g.device, dtype=dtype).to(device)
                                     for g in gens])
            else:
                gdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device=device, dtype=dtype)
        if tuple(latents.shape) != shape:
            raise ValueError(f"latents have shape {tuple(latents.shape)}, the call needs {shape}")
        if vae.config.latents_mean is None or vae.config.latents_std is None:
            raise ValueError("the VAE config carries no latents_mean / latents_std")
        mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(device, dtype)  # :448-452
        inv_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(device, dtype)  # :453-455
        vdt = vae.dtype

        video_condition = image.unsqueeze(2).to(device=device, dtype=vdt)  # :432-435, :446
        cond = retrieve_latents(vae.encode(video_condition), sample_mode="argmax")  # :464
        cond = cond.repeat(batch_size, 1, 1, 1, 1)  # :465
        cond = (cond.to(dtype) - mean) * inv_std  # :467-468

        traj = traj_tensor.to(device, dtype=vdt).unsqueeze(0).permute(0, 2, 1, 3, 4)  # :473-475 -> [1, C, F, H, W]
        traj_latents = retrieve_latents(vae.encode(traj), sample_mode="argmax")  # :478
        traj_latents = ((traj_latents - mean) * inv_std).contiguous().float()  # :481-484

        id_cond = None
        if ID_tensor is not None and ID_tensor.shape[2] != 0:  # :489
            ID_tensor = ID_tensor.to(device=device, dtype=vdt)
            parts = []
            for k in range(ID_tensor.shape[2]):  # :497-511 (the reference re-binds ID_tensor inside this loop and so
                # only survives one ID frame; each frame is encoded on its own here, as that loop intends)
                lat = retrieve_latents(vae.encode(ID_tensor[:, :, k].unsqueeze(2)), sample_mode="argmax")
                parts.append((lat.repeat(batch_size, 1, 1, 1, 1).to(dtype) - mean) * inv_std)
            id_cond = torch.cat(parts, dim=2)  # :514
            traj_latents = torch.cat([traj_latents, torch.zeros_like(id_cond)], dim=2)  # :517-518
        first_frame_mask = torch.ones(1, 1, f_lat, h_lat, w_lat, dtype=dtype, device=device)  # :529-532
        first_frame_mask[:, :, 0] = 0
        return latents, cond, traj_latents, id_cond, first_frame_mask

    # ---- :582-945 ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, image, prompt: Union[str, List[str], None] = None,
                 negative_prompt: Union[str, List[str], None] = None, traj_tensor: Optional[torch.Tensor] = None,
                 ID_tensor: Optional[torch.Tensor] = None, height: int = 480, width: int = 832, num_frames: int = 81,
                 num_inference_steps: int = 50, guidance_scale: float = 5.0, guidance_scale_2: Optional[float] = None,
                 num_videos_per_prompt: Optional[int] = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 image_embeds: Optional[torch.Tensor] = None, last_image: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "np", return_dict: bool = True,
                 attention_kwargs: Optional[Dict[str, Any]] = None, callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512,
                 fused: bool = True, cfg_parallel=None):
        """Same keywords and defaults as the reference (:582-612). Two extras: ``fused=False`` runs the reference's
        tensor-op glue around the native forward instead of the two fused kernels (same result bit for bit);
        ``cfg_parallel`` (``frameino_b200.ulysses.CfgParallel``) runs the two CFG forwards on two halves of the ranks."""
        self.check_inputs(prompt, negative_prompt, image, height, width, prompt_embeds, negative_prompt_embeds,
                          image_embeds, callback_on_step_end_tensor_inputs, guidance_scale_2)
        if image_embeds is not None or last_image is not None:
            raise NotImplementedError("image_embeds / last_image are Wan2.1 inputs (transformer.config.image_dim is "
                                      "None for Wan2.2-TI2V-5B)")
        if traj_tensor is None:
            raise ValueError("traj_tensor [F, 3, H, W] is required by the FrameINO pipeline")
        if num_frames % self.vae_scale_factor_temporal != 1:  # :707-711
            num_frames = num_frames // self.vae_scale_factor_temporal * self.vae_scale_factor_temporal + 1
        num_frames = max(num_frames, 1)
        self._guidance_scale = guidance_scale
        self._attention_kwargs = attention_kwargs
        if attention_kwargs is not None and attention_kwargs.get("scale", None) is not None:  # transformer_wan.py:463-476
            logger.warning("Passing `scale` via `attention_kwargs` when not using the PEFT backend is ineffective.")
        self._current_timestep = None
        self._interrupt = False
        device = self._execution_device
        if device.type != "cuda":
            raise RuntimeError("frameino_b200 has no CPU path: move the pipeline to a CUDA device")

        if isinstance(prompt, str):  # :724-729
            batch_size = 1
        elif isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        prompt_embeds, negative_prompt_embeds = self.encode_prompt(
            prompt, negative_prompt, self.do_classifier_free_guidance, num_videos_per_prompt, prompt_embeds,
            negative_prompt_embeds, max_sequence_length, device)  # :732-741
        tdt = self.transformer.dtype  # :745
        prompt_embeds = prompt_embeds.to(tdt)
        if negative_prompt_embeds is not None:
            negative_prompt_embeds = negative_prompt_embeds.to(tdt)
        if not self.do_classifier_free_guidance:
            negative_prompt_embeds = None  # :873 only runs the second forward under CFG

        image = self.video_processor.preprocess(image, height=height, width=width).to(device, dtype=torch.float32)  # :767
        lat, cond, traj, id_cond, mask = self.prepare_latents(
            image, traj_tensor, ID_tensor, batch_size * num_videos_per_prompt, self.vae.config.z_dim, height, width,
            num_frames, torch.float32, device, generator, latents, last_image)  # :773-787
        self._num_timesteps = num_inference_steps

        cb = None
        if callback_on_step_end is not None:  # :893-901
            state = {"prompt_embeds": prompt_embeds, "negative_prompt_embeds": negative_prompt_embeds}

            def cb(i, t, latents_now):
                self._current_timestep = t
                kwargs = {k: (latents_now if k == "latents" else state[k]) for k in callback_on_step_end_tensor_inputs}
                out = callback_on_step_end(self, i, t, kwargs) or {}
                if "prompt_embeds" in out or "negative_prompt_embeds" in out:
                    raise NotImplementedError("replacing the prompt embeddings mid-loop: the per-prompt text state is "
                                              "projected once before the loop")
                return out.get("latents", None)

        loop = wan_frameino_denoise_fused if fused else wan_frameino_denoise
        kw = dict(num_steps=num_inference_steps, guidance_scale=guidance_scale, shift=self.shift,
                  cfg_parallel=cfg_parallel, callback=cb)
        if not fused:
            kw["model_dtype"] = tdt
            if id_cond is None:
                id_cond = lat.new_zeros(lat.shape[0], lat.shape[1], 0, lat.shape[3], lat.shape[4])
        lat = loop(self.transformer, lat, cond, mask, traj, id_cond, prompt_embeds, negative_prompt_embeds, **kw)
        self._current_timestep = None

        lat = (1 - mask) * cond + mask * lat  # :914-915
        if output_type == "latent":  # :931-932
            video = lat
        else:
            z = self.vae.config.z_dim
            lat = lat.to(self.vae.dtype)  # :918
            mean = torch.tensor(self.vae.config.latents_mean).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
            inv_std = 1.0 / torch.tensor(self.vae.config.latents_std).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
            lat = lat / inv_std + mean  # :927
            video = self.vae.decode(lat, return_dict=False)[0]  # :928
            video = self.video_processor.postprocess_video(video, output_type=output_type)  # :929
        if not return_dict:
            return (video,)
        return WanPipelineOutput(frames=video)
