"""DDIM scheduler object with the surface the reference's loops touch on `model.model.scheduler`
(SURVEY.md §8b): timesteps, num_inference_steps, alphas_cumprod (CPU fp32, indexed on the host exactly like
code/models.py:89-91), final_alpha_cumprod, config.{num_train_timesteps,prediction_type}, scale_model_input,
init_noise_sigma, add_noise, step, _get_variance.

[UPSTREAM] restatement of diffusers.DDIMScheduler as configured by the checkpoints the reference loads
(beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1, set_alpha_to_one=False,
clip_sample=False).  Integer index math is exact; the device arithmetic of `step`
runs in libaedit (ae_ddim_step).
"""
from __future__ import annotations

import types
from typing import Optional

import torch


class DDIMScheduler:
    def __init__(self, beta_start: float = 0.0015, beta_end: float = 0.0195, num_train_timesteps: int = 1000,
                 prediction_type: str = "epsilon", steps_offset: int = 1):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                                            steps_offset=steps_offset, beta_start=beta_start, beta_end=beta_end)
        self.num_inference_steps: Optional[int] = None
        self.timesteps: Optional[torch.Tensor] = None
        self._table = None   # libaedit scheduler table, built lazily by the wrapper

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps larger than num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps, dtype=torch.int64) * step_ratio).flip(0) + self.config.steps_offset
        self.timesteps_cpu = ts
        self.timesteps = ts.to(device) if device is not None else ts
        self._table = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _get_variance(self, timestep, prev_timestep):
        a_t = self.alphas_cumprod[int(timestep)]
        a_p = self.alphas_cumprod[int(prev_timestep)] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        so = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        so = so.flatten()
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            so = so.unsqueeze(-1)
        return sa * original_samples + so * noise
