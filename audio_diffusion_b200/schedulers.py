"""`DDPMScheduler` / `DDIMScheduler` — drop-ins for the diffusers classes at the call sites of the reference:
`set_timesteps`, `timesteps`, `step(...)["prev_sample"]`, `add_noise`, `alphas_cumprod`,
`final_alpha_cumprod`, `num_inference_steps`, `config.num_train_timesteps`
(audiodiffusion/pipeline_audio_diffusion.py:115,150,157,159,165-179,221-234; scripts/train_unet.py:161-164,250).

Coefficients are computed exactly as diffusers does (fp32 0-d torch tensors on the host).  `step()` is the
reference-compatible elementwise path (torch ops on whatever device the tensors live on); `step_coef()` hands
the same scalars to the fused U-Net output kernel (`UNet2DModel.forward_step`), which is what
`AudioDiffusionPipeline` uses so that no separate elementwise launch runs per step.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np
import torch

from ._lib import StepCoefC


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class SchedulerOutput(dict):
    def __init__(self, prev_sample, pred_original_sample=None):
        super().__init__(prev_sample=prev_sample, pred_original_sample=pred_original_sample)
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


def randn_tensor(shape, generator, device, dtype=torch.float32) -> torch.Tensor:
    """diffusers.utils.torch_utils.randn_tensor as the schedulers use it ([3P-recall] 0.24): a CPU generator with a CUDA
    target draws on the CPU and moves the result (so CPU-seeded runs reproduce across devices)."""
    device = torch.device(device)
    if generator is not None and generator.device.type != device.type and generator.device.type == "cpu":
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


class _SchedulerBase:
    config_name = "scheduler_config.json"
    _class_name = "SchedulerBase"

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 beta_schedule: str = "linear", clip_sample: bool = True, clip_sample_range: float = 1.0,
                 prediction_type: str = "epsilon", timestep_spacing: str = "leading", steps_offset: int = 0, **extra):
        if beta_schedule != "linear" or prediction_type != "epsilon" or timestep_spacing != "leading":
            raise ValueError("b200 schedulers implement the reference defaults: linear betas, epsilon, leading")
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, clip_sample=clip_sample, clip_sample_range=clip_sample_range,
                           prediction_type=prediction_type, timestep_spacing=timestep_spacing,
                           steps_offset=steps_offset, _class_name=self._class_name, **extra)
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    # -- diffusers-compatible API ----------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps) -> torch.Tensor:
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = torch.as_tensor(timesteps).to(original_samples.device)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        while sa.ndim < original_samples.ndim:
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def __len__(self):
        return self.config.num_train_timesteps

    # -- persistence (diffusers directory layout: scheduler/scheduler_config.json) -------------------
    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, self.config_name), "w") as f:
            json.dump(dict(self.config), f, indent=2)

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        with open(os.path.join(path, cls.config_name)) as f:
            cfg = json.load(f)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        cfg.update(kw)
        known = {"num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "clip_sample", "clip_sample_range",
                 "prediction_type", "timestep_spacing", "steps_offset", "set_alpha_to_one", "variance_type"}
        return cls(**{k: v for k, v in cfg.items() if k in known})

    def _noise(self, like: torch.Tensor, generator):
        return randn_tensor(like.shape, generator, like.device, like.dtype)


class DDPMScheduler(_SchedulerBase):
    _class_name = "DDPMScheduler"

    def __init__(self, num_train_timesteps: int = 1000, variance_type: str = "fixed_small", **kw):
        if variance_type != "fixed_small":
            raise ValueError("only variance_type='fixed_small' (the reference default) is implemented")
        super().__init__(num_train_timesteps, variance_type=variance_type, **kw)

    def _scalars(self, t: int):
        n = self.num_inference_steps or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        c_x0 = (a_prev ** 0.5 * cur_b) / b_t
        c_xt = cur_a ** 0.5 * b_prev / b_t
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_b, min=1e-20)
        return a_t, b_t, c_x0, c_xt, var

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        t = int(timestep)
        a_t, b_t, c_x0, c_xt, var = self._scalars(t)
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        prev = c_x0 * x0 + c_xt * sample
        if t > 0:
            prev = prev + (var ** 0.5) * self._noise(model_output, generator)
        return SchedulerOutput(prev, x0) if return_dict else (prev,)

    def needs_noise(self, timestep, eta: float = 0.0) -> bool:
        return int(timestep) > 0

    def step_coef(self, timestep, eta: float = 0.0) -> StepCoefC:
        t = int(timestep)
        a_t, b_t, c_x0, c_xt, var = self._scalars(t)
        c = StepCoefC()
        c.sqrt_1m_at = float(b_t ** 0.5)
        c.inv_sqrt_at = float(1.0 / a_t ** 0.5)
        c.clip = float(self.config.clip_sample_range)
        c.do_clip = 1 if self.config.clip_sample else 0
        c.c_x0, c.c_xt, c.c_eps = float(c_x0), float(c_xt), 0.0
        c.c_z = float(var ** 0.5) if t > 0 else 0.0
        return c


class DDIMScheduler(_SchedulerBase):
    _class_name = "DDIMScheduler"

    def __init__(self, num_train_timesteps: int = 1000, set_alpha_to_one: bool = True, **kw):
        super().__init__(num_train_timesteps, set_alpha_to_one=set_alpha_to_one, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def _scalars(self, t: int, eta: float):
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        var = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * var ** 0.5
        return a_t, a_prev, b_t, std

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if use_clipped_model_output:
            raise ValueError("use_clipped_model_output=True is not used by the reference and is not implemented")
        t = int(timestep)
        a_t, a_prev, b_t, std = self._scalars(t, eta)
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            z = variance_noise if variance_noise is not None else self._noise(model_output, generator)
            prev = prev + std * z
        return SchedulerOutput(prev, x0) if return_dict else (prev,)

    def needs_noise(self, timestep, eta: float = 0.0) -> bool:
        return eta > 0

    def invert_step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor) -> torch.Tensor:
        """One step of deterministic DDIM inversion, x_{t-Δ} -> x_t (audiodiffusion/pipeline_audio_diffusion.py:229-240):
        remove the ε-direction at the lower noise level, rescale to x0, then re-noise to level t with the same ε."""
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (sample - (1 - a_prev) ** 0.5 * model_output) * a_prev ** (-0.5)
        return x0 * a_t ** 0.5 + (1 - a_t) ** 0.5 * model_output

    def step_coef(self, timestep, eta: float = 0.0) -> StepCoefC:
        t = int(timestep)
        a_t, a_prev, b_t, std = self._scalars(t, eta)
        c = StepCoefC()
        c.sqrt_1m_at = float(b_t ** 0.5)
        c.inv_sqrt_at = float(1.0 / a_t ** 0.5)
        c.clip = float(self.config.clip_sample_range)
        c.do_clip = 1 if self.config.clip_sample else 0
        c.c_x0 = float(a_prev ** 0.5)
        c.c_xt = 0.0
        c.c_eps = float((1 - a_prev - std ** 2) ** 0.5)
        c.c_z = float(std) if eta > 0 else 0.0
        return c
