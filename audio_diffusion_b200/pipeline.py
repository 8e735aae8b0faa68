"""`AudioDiffusionPipeline` — same call surface as audiodiffusion/pipeline_audio_diffusion.py:39-258
(`__call__`, `encode`, `slerp`, `get_default_steps`) plus the minimal `DiffusionPipeline` base the reference
relies on (`register_modules`, `device`, `progress_bar`, `to`, `save_pretrained`, `from_pretrained`).

What differs underneath (and only underneath):
  * the denoise loop calls `UNet2DModel.forward_step`, i.e. U-Net forward with the DDPM/DDIM update fused into
    the output kernel (reference: two separate calls at :163 and :165-179);
  * float -> uint8 conversion runs on the GPU and is bit-exact with :192-194;
  * `Mel.image_to_audio` is run batched on the GPU instead of one image at a time on the CPU (:201).
RNG stays in PyTorch: initial noise and per-step noise are drawn from the caller's `torch.Generator` with the
same calls, shapes and order as the reference, so seeds reproduce the reference stream (:120-130, :171, :178).
"""
from __future__ import annotations

import json
import os
from math import acos, sin
from typing import List, Optional, Union

import numpy as np
import torch
from PIL import Image

from . import _lib
from .mel import Mel
from .schedulers import DDIMScheduler, DDPMScheduler, randn_tensor
from .unet import UNet2DModel
from .vae import AutoencoderKL


class BaseOutput(dict):
    def __init__(self, **kw):
        super().__init__(**kw)
        for k, v in kw.items():
            setattr(self, k, v)


class AudioPipelineOutput(BaseOutput):
    def __init__(self, audios):
        super().__init__(audios=audios)


class ImagePipelineOutput(BaseOutput):
    def __init__(self, images):
        super().__init__(images=images)


class DiffusionPipeline:
    config_name = "model_index.json"
    _optional_components: List[str] = []

    def __init__(self):
        self._modules_ = {}
        self._progress = None

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            self._modules_[k] = v
            setattr(self, k, v)

    @property
    def device(self) -> torch.device:
        for m in self._modules_.values():
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device("cpu")

    def to(self, device):
        for k, m in self._modules_.items():
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def set_progress_bar_config(self, **kw):
        self._progress_kw = kw

    def progress_bar(self, iterable):
        kw = getattr(self, "_progress_kw", {})
        if kw.get("disable"):
            return iterable
        try:
            from tqdm.auto import tqdm
            return tqdm(iterable, **kw)
        except Exception:
            return iterable

    # -- diffusers directory layout: model_index.json + one sub-directory per module -----------------------
    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        index = {"_class_name": type(self).__name__}
        for k, m in self._modules_.items():
            if m is None:
                index[k] = [None, None]
                continue
            index[k] = ["audio_diffusion_b200", type(m).__name__]
            sub = os.path.join(path, k)
            if isinstance(m, (UNet2DModel, AutoencoderKL)):
                os.makedirs(sub, exist_ok=True)
                with open(os.path.join(sub, "config.json"), "w") as f:
                    json.dump({kk: vv for kk, vv in m.config.items()}, f, indent=2)
                from safetensors.torch import save_file
                save_file({kk: vv.detach().cpu().contiguous() for kk, vv in m.state_dict().items()},
                          os.path.join(sub, "diffusion_pytorch_model.safetensors"))
            else:
                m.save_pretrained(sub)
        with open(os.path.join(path, self.config_name), "w") as f:
            json.dump(index, f, indent=2)

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        with open(os.path.join(path, cls.config_name)) as f:
            index = json.load(f)
        mods = {}
        for k, v in index.items():
            if k.startswith("_") or not isinstance(v, list):
                continue
            sub = os.path.join(path, k)
            cname = v[1]
            if cname is None or not os.path.isdir(sub):
                mods[k] = None
            elif cname in ("UNet2DModel",):
                mods[k] = load_unet(sub)
            elif cname == "AutoencoderKL":
                mods[k] = load_vae(sub)
            elif cname == "DDPMScheduler":
                mods[k] = DDPMScheduler.from_pretrained(sub)
            elif cname == "DDIMScheduler":
                mods[k] = DDIMScheduler.from_pretrained(sub)
            elif cname == "Mel":
                mods[k] = Mel.from_pretrained(sub)
            else:
                raise ValueError(f"from_pretrained: component {k} of class {cname} is not supported by the b200 engine")
        mods.setdefault("vqvae", None)
        return cls(**mods)


def load_unet(sub: str) -> UNet2DModel:
    """Load `unet/config.json` + `diffusion_pytorch_model.{safetensors,bin}` (diffusers layout), including the
    deprecated attention key names (query/key/value/proj_attn) of older hub files."""
    with open(os.path.join(sub, "config.json")) as f:
        cfg = json.load(f)
    keep = ("sample_size", "in_channels", "out_channels", "down_block_types", "up_block_types", "block_out_channels",
            "layers_per_block", "attention_head_dim", "norm_num_groups", "norm_eps")
    model = UNet2DModel(**{k: cfg[k] for k in keep if k in cfg})
    model.load_state_dict(_load_weights(sub))
    return model


def _load_weights(sub: str):
    st = os.path.join(sub, "diffusion_pytorch_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(sub, "diffusion_pytorch_model.bin"), map_location="cpu")
    ren = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}
    fixed = {}
    for k, v in sd.items():
        for a, b in ren.items():
            k = k.replace(a, b)
        fixed[k] = v.to(torch.float32)
    return fixed


def load_vae(sub: str) -> AutoencoderKL:
    """`vqvae/` component of a latent pipeline (diffusers AutoencoderKL layout)."""
    return AutoencoderKL.from_pretrained(sub)


class AudioDiffusionPipeline(DiffusionPipeline):
    _optional_components = ["vqvae"]

    def __init__(self, vqvae, unet: UNet2DModel, mel: Mel, scheduler: Union[DDIMScheduler, DDPMScheduler]):
        super().__init__()
        self.register_modules(unet=unet, scheduler=scheduler, mel=mel, vqvae=vqvae)

    def get_default_steps(self) -> int:
        return 50 if isinstance(self.scheduler, DDIMScheduler) else 1000

    @torch.no_grad()
    def __call__(
        self,
        batch_size: int = 1,
        audio_file: str = None,
        raw_audio: np.ndarray = None,
        slice: int = 0,
        start_step: int = 0,
        steps: int = None,
        generator: torch.Generator = None,
        mask_start_secs: float = 0,
        mask_end_secs: float = 0,
        step_generator: torch.Generator = None,
        eta: float = 0,
        noise: torch.Tensor = None,
        encoding: torch.Tensor = None,
        return_dict=True,
        return_audio: bool = True,
    ):
        if encoding is not None:
            raise NotImplementedError("conditional generation (UNet2DConditionModel) is outside the b200 hot path")
        steps = steps or self.get_default_steps()
        self.scheduler.set_timesteps(steps)
        step_generator = step_generator or generator
        if type(self.unet.sample_size) == int:  # backwards compatibility, as the reference (:118-119)
            self.unet.sample_size = (self.unet.sample_size, self.unet.sample_size)
        device = self.device
        if noise is None:
            noise = torch.randn(
                (batch_size, self.unet.in_channels, self.unet.sample_size[0], self.unet.sample_size[1]),
                generator=generator, device=device)
        images = noise.to(device=device, dtype=torch.float32).clone()
        mask = None
        mask_start = mask_end = 0

        if audio_file is not None or raw_audio is not None:
            self.mel.load_audio(audio_file, raw_audio)
            input_image = self.mel.audio_slice_to_image(slice)
            input_image = np.frombuffer(input_image.tobytes(), dtype="uint8").reshape(
                (input_image.height, input_image.width))
            input_image = (input_image / 255) * 2 - 1
            input_images = torch.tensor(input_image[np.newaxis, :, :], dtype=torch.float).to(device)
            if self.vqvae is not None:  # latent audio diffusion (:143-147)
                input_images = self.vqvae.encode(torch.unsqueeze(input_images, 0)).latent_dist.sample(
                    generator=generator)[0]
                input_images = 0.18215 * input_images
            if start_step > 0:
                images[0, 0] = self.scheduler.add_noise(input_images, noise, self.scheduler.timesteps[start_step - 1])
            pixels_per_second = (
                self.unet.sample_size[1] * self.mel.get_sample_rate() / self.mel.x_res / self.mel.hop_length)
            mask_start = int(mask_start_secs * pixels_per_second)
            mask_end = int(mask_end_secs * pixels_per_second)
            # The reference's `images = noise` ALIASES the two tensors (:131), so the assignment above also rewrites
            # noise[0, 0] and the mask below is built from that modified noise. `images` (a private copy here, because the
            # fused step updates it in place) holds exactly those values at this point.
            mask = self.scheduler.add_noise(input_images, images, torch.tensor(self.scheduler.timesteps[start_step:]))

        fused = isinstance(self.unet, UNet2DModel) and hasattr(self.scheduler, "step_coef")
        is_ddim = isinstance(self.scheduler, DDIMScheduler)
        for step, t in enumerate(self.progress_bar(self.scheduler.timesteps[start_step:])):
            if fused:
                z = None
                if self.scheduler.needs_noise(t, eta):
                    z = randn_tensor(images.shape, step_generator, images.device, images.dtype)  # as scheduler.step draws it
                images = self.unet.forward_step(images, t, self.scheduler.step_coef(t, eta), noise=z, out=images)
            else:
                model_output = self.unet(images, t)["sample"]
                if is_ddim:
                    images = self.scheduler.step(model_output=model_output, timestep=t, sample=images, eta=eta,
                                                 generator=step_generator)["prev_sample"]
                else:
                    images = self.scheduler.step(model_output=model_output, timestep=t, sample=images,
                                                 generator=step_generator)["prev_sample"]
            if mask is not None:
                if mask_start > 0:
                    images[:, :, :, :mask_start] = mask[:, step, :, :mask_start]
                if mask_end > 0:
                    images[:, :, :, -mask_end:] = mask[:, step, :, -mask_end:]

        if self.vqvae is not None:
            # 0.18215 was scaling factor used in training to ensure unit variance (:187-190)
            images = 1 / 0.18215 * images
            images = self.vqvae.decode(images)["sample"]

        u8 = self.images_to_u8(images)                      # (B, C, H, W) uint8 on the device
        host = u8.permute(0, 2, 3, 1).cpu().numpy()
        pil = list(map(lambda _: Image.fromarray(_[:, :, 0]), host) if host.shape[3] == 1
                   else map(lambda _: Image.fromarray(_, mode="RGB").convert("L"), host))
        if not return_audio:
            return pil
        if host.shape[3] == 1 and hasattr(self.mel, "images_to_audio"):
            audios = list(self.mel.images_to_audio(u8[:, 0]))          # batched Griffin-Lim (the engine's Mel)
        else:                                                          # any object with the reference Mel's surface (:201)
            audios = list(map(lambda _: self.mel.image_to_audio(_), pil))
        if not return_dict:
            return pil, (self.mel.get_sample_rate(), audios)
        return BaseOutput(**AudioPipelineOutput(np.array(audios)[:, np.newaxis, :]), **ImagePipelineOutput(pil))

    @staticmethod
    def images_to_u8(images: torch.Tensor) -> torch.Tensor:
        """`(images/2+0.5).clamp(0,1)*255 -> round -> uint8` (:192-194), bit-exact, on the device."""
        x = images.to(torch.float32).contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().b200ad_sample_to_u8(x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()))
        return out

    @torch.no_grad()
    def encode(self, images: List[Image.Image], steps: int = 50) -> torch.Tensor:
        """DDIM inversion (:207-242): recover the noise that generates `images`."""
        assert isinstance(self.scheduler, DDIMScheduler)
        self.scheduler.set_timesteps(steps)
        sample = np.array(
            [np.frombuffer(image.tobytes(), dtype="uint8").reshape((1, image.height, image.width)) for image in images])
        sample = (sample / 255) * 2 - 1
        sample = torch.Tensor(sample).to(self.device)
        for t in self.progress_bar(torch.flip(self.scheduler.timesteps, (0,))):
            prev_timestep = t - self.scheduler.config.num_train_timesteps // self.scheduler.num_inference_steps
            alpha_prod_t = self.scheduler.alphas_cumprod[t]
            alpha_prod_t_prev = (self.scheduler.alphas_cumprod[prev_timestep] if prev_timestep >= 0
                                 else self.scheduler.final_alpha_cumprod)
            beta_prod_t = 1 - alpha_prod_t
            model_output = self.unet(sample, t)["sample"]
            pred_sample_direction = (1 - alpha_prod_t_prev) ** (0.5) * model_output
            sample = (sample - pred_sample_direction) * alpha_prod_t_prev ** (-0.5)
            sample = sample * alpha_prod_t ** (0.5) + beta_prod_t ** (0.5) * model_output
        return sample

    @staticmethod
    def slerp(x0: torch.Tensor, x1: torch.Tensor, alpha: float) -> torch.Tensor:
        """Spherical linear interpolation (:244-258)."""
        theta = acos(torch.dot(torch.flatten(x0), torch.flatten(x1)) / torch.norm(x0) / torch.norm(x1))
        return sin((1 - alpha) * theta) * x0 / sin(theta) + sin(alpha * theta) * x1 / sin(theta)
