"""Minimal `diffusers` import surface for the unchanged teticio/audio-diffusion sources, backed by audio_diffusion_b200.

Only the names the reference imports exist (pipeline_audio_diffusion.py:24-33, mel.py:22-23, train_unet.py:17-20,
audio_encoder.py:3, audio_to_images.py:10)."""
from audio_diffusion_b200.mel import Mel
from audio_diffusion_b200.pipeline import (AudioPipelineOutput, BaseOutput, DiffusionPipeline, ImagePipelineOutput)
from audio_diffusion_b200.schedulers import DDIMScheduler, DDPMScheduler
from audio_diffusion_b200.unet import UNet2DModel
from audio_diffusion_b200.vae import AutoencoderKL

from .configuration_utils import ConfigMixin, register_to_config  # noqa: F401

__version__ = "0.24.0+b200shim"


class ModelMixin:  # audio_encoder.py:3 (conditional path, out of scope) — import surface only
    pass


class _NotBuilt:
    _what = "component"

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}: the {self._what} is outside the b200 hot path (SURVEY §8)")

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise EnvironmentError(f"{cls.__name__} is not available in the b200 engine")


class UNet2DConditionModel(_NotBuilt):  # isinstance() discriminator at pipeline_audio_diffusion.py:160
    _what = "conditional U-Net"

