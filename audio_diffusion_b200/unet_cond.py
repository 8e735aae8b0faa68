"""`UNet2DConditionModel` — drop-in for `diffusers.UNet2DConditionModel` as the reference constructs it
(scripts/train_unet.py:139-159) and calls it (`unet(sample, t, encoding)["sample"]`,
audiodiffusion/pipeline_audio_diffusion.py:160-161; scripts/train_unet.py:255): the conditional audio-diffusion model whose
cross-attention reads the 100-d audio encodings of `audiodiffusion/audio_encoder.py:62-107`.

Same constructor kwargs and diffusers state-dict keys (`down_blocks.i.attentions.j.transformer_blocks.0.attn1.to_q.weight`, ...).
Inference runs in libb200ad.so: every projection / linear of the transformer blocks on the tcgen05 conv kernel, self-attention
(8 heads, head_dim = channels / 8) on a flash-style tensor-core kernel, cross-attention against the ONE encoder token as a
per-sample vector folded into the attn1 output projection.  Limits: encoder sequence length 1 (what the reference's
AudioEncoder produces: (B, 1, 100)); no backward (training the conditional model is not built).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import MAX_BLOCKS, StepCoefC, UNetConfigC
from .unet import UNet2DModel, UNet2DOutput, _Cfg


class UNet2DConditionModel(UNet2DModel):
    is_conditional = True

    def __init__(
        self,
        sample_size: Optional[Union[int, Tuple[int, int]]] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        center_input_sample: bool = False,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Sequence[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
        mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
        up_block_types: Sequence[str] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
        only_cross_attention: bool = False,
        block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
        layers_per_block: int = 2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        dropout: float = 0.0,
        act_fn: str = "silu",
        norm_num_groups: int = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim: int = 1280,
        transformer_layers_per_block: int = 1,
        attention_head_dim: int = 8,
        num_attention_heads: Optional[int] = None,
        dual_cross_attention: bool = False,
        use_linear_projection: bool = False,
        class_embed_type: Optional[str] = None,
        upcast_attention: bool = False,
        resnet_time_scale_shift: str = "default",
        time_embedding_type: str = "positional",
        seed: Optional[int] = None,
    ):
        torch.nn.Module.__init__(self)
        bad = []
        if center_input_sample or not flip_sin_to_cos or freq_shift != 0: bad.append("input / timestep embedding options")
        if mid_block_type != "UNetMidBlock2DCrossAttn": bad.append("mid_block_type")
        if only_cross_attention or dual_cross_attention or use_linear_projection or upcast_attention: bad.append("attention options")
        if downsample_padding != 1 or mid_block_scale_factor != 1 or dropout: bad.append("padding / scale / dropout")
        if act_fn != "silu" or resnet_time_scale_shift != "default" or time_embedding_type != "positional": bad.append("act / shift / time")
        if transformer_layers_per_block != 1 or class_embed_type is not None: bad.append("transformer depth / class embedding")
        heads = num_attention_heads if num_attention_heads is not None else attention_head_dim   # diffusers 0.24 naming quirk
        if heads != 8: bad.append("number of attention heads != 8")
        if len(block_out_channels) > MAX_BLOCKS or len(down_block_types) != len(block_out_channels): bad.append("blocks")
        for t in down_block_types:
            if t not in ("DownBlock2D", "CrossAttnDownBlock2D"): bad.append(t)
        for t in up_block_types:
            if t not in ("UpBlock2D", "CrossAttnUpBlock2D"): bad.append(t)
        if bad:
            raise ValueError(f"UNet2DConditionModel(b200): unsupported configuration: {bad}")
        self.sample_size = sample_size
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.config = _Cfg(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), mid_block_type=mid_block_type, up_block_types=tuple(up_block_types),
            block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
            cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps, _class_name="UNet2DConditionModel")
        c = UNetConfigC()
        c.in_channels, c.out_channels = in_channels, out_channels
        c.layers_per_block, c.num_blocks = layers_per_block, len(block_out_channels)
        for i, v in enumerate(block_out_channels):
            c.block_out_channels[i] = int(v)
            c.down_cross[i] = 1 if down_block_types[i] == "CrossAttnDownBlock2D" else 0
            c.up_cross[i] = 1 if up_block_types[i] == "CrossAttnUpBlock2D" else 0
        c.norm_num_groups, c.norm_eps = norm_num_groups, norm_eps
        c.attention_head_dim = heads
        c.cross_attention_dim = cross_attention_dim
        self._init_engine(c, seed)

    def _encoding(self, enc: torch.Tensor, n: int, dev) -> torch.Tensor:
        if enc is None:
            raise ValueError("UNet2DConditionModel needs encoder_hidden_states (B, seq, cross_attention_dim)")
        e = enc.to(device=dev, dtype=torch.float32)
        if e.ndim == 2:
            e = e[:, None, :]
        if e.ndim != 3 or e.shape[0] != n or e.shape[2] != self.config.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states must be ({n}, seq, {self.config.cross_attention_dim}), got {tuple(enc.shape)}")
        if e.shape[1] != 1:
            raise NotImplementedError("UNet2DConditionModel(b200): encoder sequence length 1 only (audio_encoder.py encodings)")
        return e.contiguous()

    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor = None, return_dict: bool = True):
        """ε = unet(sample, timestep, encoding)["sample"] — pipeline_audio_diffusion.py:161."""
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("UNet2DConditionModel(b200): inference only — call .eval() / torch.no_grad()")
        x = self._check_input(sample)
        n, _, hh, ww = x.shape
        with torch.cuda.device(x.device):
            self._set_training_mode(False)
            self._ensure_bound(n, hh, ww)
            t = self._timesteps(timestep, n, x.device)
            e = self._encoding(encoder_hidden_states, n, x.device)
            out = torch.empty((n, self.out_channels, hh, ww), dtype=torch.float32, device=x.device)
            L = _lib.lib()
            _lib.check(L.b200ad_unet_set_encoding(self._h, e.data_ptr(), e.shape[1]))
            _lib.check(L.b200ad_unet_forward(self._h, x.data_ptr(), t.data_ptr(), out.data_ptr(), _lib.stream_ptr()))
        return UNet2DOutput(out) if return_dict else (out,)

    @torch.no_grad()
    def forward_step(self, sample: torch.Tensor, timestep, coef: StepCoefC, encoder_hidden_states: torch.Tensor = None,
                     noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, want_eps: bool = False):
        """Fused `scheduler.step(unet(sample, t, encoding), t, sample)["prev_sample"]` (pipeline_audio_diffusion.py:161-179)."""
        x = self._check_input(sample)
        n, _, hh, ww = x.shape
        with torch.cuda.device(x.device):
            self._set_training_mode(False)
            self._ensure_bound(n, hh, ww)
            t = self._timesteps(timestep, n, x.device)
            e = self._encoding(encoder_hidden_states, n, x.device)
            if out is None:
                out = torch.empty_like(x)
            eps = torch.empty_like(x) if want_eps else None
            z = noise.to(torch.float32).contiguous() if noise is not None else None
            L = _lib.lib()
            _lib.check(L.b200ad_unet_set_encoding(self._h, e.data_ptr(), e.shape[1]))
            _lib.check(L.b200ad_unet_forward_step(
                self._h, x.data_ptr(), t.data_ptr(), z.data_ptr() if z is not None else None, C.byref(coef),
                out.data_ptr(), eps.data_ptr() if eps is not None else None, _lib.stream_ptr()))
        return (out, eps) if want_eps else out


def load_cond_unet(sub: str) -> UNet2DConditionModel:
    return UNet2DConditionModel.from_pretrained(sub)
