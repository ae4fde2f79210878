"""Flow-match Euler with a static shift — the stepper the parity tests use on BOTH sides (the scheduler classes are
upstream-only; config/train_wan_motion_FrameINO.yaml:43-50 gives shift 5.0). Same arithmetic as
frameino_b200.sampling.flow_match_sigmas + the Euler update, kept in fp32."""
import torch

from ..configuration_utils import FrozenDict


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0):
        self.config = FrozenDict(num_train_timesteps=num_train_timesteps, shift=shift)
        self.timesteps = None
        self.sigmas = None
        self._i = 0

    def set_timesteps(self, num_inference_steps, device=None):
        s = torch.linspace(1.0, 1.0 / self.config.num_train_timesteps, num_inference_steps, dtype=torch.float32)
        shift = self.config.shift
        s = shift * s / (1.0 + (shift - 1.0) * s)
        self.sigmas = torch.cat([s, s.new_zeros(1)]).to(device)
        self.timesteps = (s * self.config.num_train_timesteps).to(device)
        self._i = 0

    def step(self, model_output, timestep, sample, return_dict=True):
        i = self._i
        prev = sample.float() + (self.sigmas[i + 1] - self.sigmas[i]) * model_output.float()
        self._i += 1
        return (prev,) if not return_dict else FrozenDict(prev_sample=prev)
