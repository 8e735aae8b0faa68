"""TEST-ONLY overlay of the `diffusers` import surface (audio_diffusion_b200/compat/diffusers): identical, except that
`UNet2DModel` is an autograd-capable CPU module backed by the oracle (oracle/unet_oracle.py), so that the UNCHANGED
reference scripts/train_unet.py can be driven end to end in the GPU-less container (the product's UNet2DModel has no CPU
path by design).  Only tests/ may put this directory on sys.path."""
import os

import torch
from torch import nn

import audio_diffusion_b200.compat.diffusers as _real
from audio_diffusion_b200.compat.diffusers import *  # noqa: F401,F403
from audio_diffusion_b200.compat.diffusers import (AutoencoderKL, ConfigMixin, DDIMScheduler, DDPMScheduler,  # noqa: F401
                                                   DiffusionPipeline, Mel, ModelMixin, UNet2DConditionModel)

__path__.append(os.path.dirname(_real.__file__))      # diffusers.optimization, .training_utils, .pipelines... resolve there
__version__ = _real.__version__


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class UNet2DModel(nn.Module):
    """Same ctor kwargs / state-dict keys / call convention as the product's UNet2DModel; forward = the fp32 oracle."""

    def __init__(self, sample_size=None, in_channels=1, out_channels=1, layers_per_block=2,
                 block_out_channels=(128, 128, 256, 256, 512, 512), down_block_types=(), up_block_types=(), **kw):
        super().__init__()
        from oracle.unet_oracle import UNetConfig, init_weights
        size = tuple(sample_size) if isinstance(sample_size, (tuple, list)) else (sample_size, sample_size)
        self.ocfg = UNetConfig(sample_size=size, in_channels=in_channels, out_channels=out_channels,
                               layers_per_block=layers_per_block, block_out_channels=tuple(block_out_channels),
                               down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types))
        self.sample_size, self.in_channels, self.out_channels = sample_size, in_channels, out_channels
        self.config = _Cfg(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                           layers_per_block=layers_per_block, block_out_channels=tuple(block_out_channels),
                           down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                           _class_name="UNet2DModel")
        self._names = []
        for k, v in init_weights(self.ocfg, seed=0).items():
            self._names.append(k)
            self.register_parameter(k.replace(".", "__"), nn.Parameter(v))

    @classmethod
    def from_pretrained(cls, path, **_unused):
        from audio_diffusion_b200.hub_io import model_from_dir
        return model_from_dir(cls, path)

    def save_pretrained(self, path, **_unused):
        from audio_diffusion_b200.hub_io import save_model
        save_model(self, path)

    def _w(self):
        return {k: getattr(self, k.replace(".", "__")) for k in self._names}

    def state_dict(self, *a, **k):
        return {k_: v.detach() for k_, v in self._w().items()}

    def load_state_dict(self, sd, strict=True):
        with torch.no_grad():
            for k_, p in self._w().items():
                p.copy_(sd[k_])

    def forward(self, sample, timestep):
        from oracle.unet_oracle import unet_forward
        return {"sample": unet_forward(self._w(), self.ocfg, sample, timestep)}
