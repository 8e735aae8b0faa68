"""Restatement of diffusers==0.24.0 DDPMScheduler / DDIMScheduler (defaults) in plain torch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (diffusers absent).
Call sites followed: audiodiffusion/pipeline_audio_diffusion.py:115 (set_timesteps), :150,:157
(add_noise), :165-179 (step), :221-234 (alphas_cumprod / final_alpha_cumprod); scripts/train_unet.py:161-164,250.

[3P-recall] defaults: num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
variance_type="fixed_small", clip_sample=True, clip_sample_range=1, prediction_type="epsilon",
timestep_spacing="leading", steps_offset=0; DDIM: set_alpha_to_one=True.
"""
from __future__ import annotations

import numpy as np
import torch


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class OracleDDPM:
    def __init__(self, num_train_timesteps: int = 1000, beta_start=1e-4, beta_end=0.02):
        self.config = _Cfg(num_train_timesteps=num_train_timesteps)
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, n: int):
        T = self.config.num_train_timesteps
        self.num_inference_steps = n
        ratio = T // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def _prev(self, t):
        n = self.num_inference_steps or self.config.num_train_timesteps
        return t - self.config.num_train_timesteps // n

    def step(self, model_output, timestep, sample, generator=None):
        t = int(timestep)
        prev_t = self._prev(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        x0 = x0.clamp(-1.0, 1.0)
        c_x0 = (a_prev ** 0.5 * cur_b) / b_t
        c_xt = cur_a ** 0.5 * b_prev / b_t
        prev = c_x0 * x0 + c_xt * sample
        if t > 0:
            z = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                            dtype=model_output.dtype)
            var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_b, min=1e-20)
            prev = prev + (var ** 0.5) * z
        return {"prev_sample": prev, "pred_original_sample": x0}

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original.device, dtype=original.dtype)
        timesteps = torch.as_tensor(timesteps).to(original.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.ndim < original.ndim:
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original + sb * noise


class OracleDDIM(OracleDDPM):
    def __init__(self, num_train_timesteps: int = 1000, **kw):
        super().__init__(num_train_timesteps, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0)

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        x0 = x0.clamp(-1.0, 1.0)
        var = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * var ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output  # unclipped epsilon
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            z = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                            dtype=model_output.dtype)
            prev = prev + std * z
        return {"prev_sample": prev, "pred_original_sample": x0}
