"""frameino_b200 — B200-native (sm_100a) denoise-step forward for FrameINO's Wan2.2 / CogVideoX transformers.

Hot path only: the per-step forward of the motion/ID-conditioned DiT behind the reference's
``WanTransformer3DModel.forward`` / ``AttnProcessor`` surface. Hand-written CUDA behind a C ABI
(``include/frameino_b200.h``, ``frameino_b200/csrc``); PyTorch is plumbing (memory, streams, torch.distributed).
"""

__version__ = "0.1.0"
