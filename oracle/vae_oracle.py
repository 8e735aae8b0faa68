"""Plain-PyTorch fp32 restatement of diffusers==0.24.0 `AutoencoderKL` (Encoder / Decoder / DiagonalGaussianDistribution)
for the shape of config/ldm_autoencoder_kl.yaml:18-28 (ch 128, ch_mult [1,2,4,4], 2 res blocks, z_channels 1,
1-channel in/out, mid-block single-head attention).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (diffusers absent).
Call sites served: audiodiffusion/pipeline_audio_diffusion.py:143-147 (`vqvae.encode(x).latent_dist.sample(generator)`),
:187-190 (`vqvae.decode(z)["sample"]`), scripts/train_unet.py:99-104,230-235.  State-dict keys as produced by
audiodiffusion/utils.py:156-303 (`convert_ldm_to_hf_vae`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 1
    out_channels: int = 1
    latent_channels: int = 1
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    # [3P-recall] Encoder/Decoder resnets and norms use eps=1e-6; mid-block attention_head_dim = block_out_channels[-1]


EPS = 1e-6


def param_shapes(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}
    boc = cfg.block_out_channels

    def conv(n, ci, co, k):
        sh[n + ".weight"] = (co, ci, k, k)
        sh[n + ".bias"] = (co,)

    def lin(n, ci, co):
        sh[n + ".weight"] = (co, ci)
        sh[n + ".bias"] = (co,)

    def gn(n, c):
        sh[n + ".weight"] = (c,)
        sh[n + ".bias"] = (c,)

    def resnet(n, ci, co):
        gn(n + ".norm1", ci)
        conv(n + ".conv1", ci, co, 3)
        gn(n + ".norm2", co)
        conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".conv_shortcut", ci, co, 1)

    def attn(n, c):
        gn(n + ".group_norm", c)
        for p in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(n + "." + p, c, c)

    def mid(n, c):
        resnet(n + ".resnets.0", c, c)
        attn(n + ".attentions.0", c)
        resnet(n + ".resnets.1", c, c)

    # encoder
    conv("encoder.conv_in", cfg.in_channels, boc[0], 3)
    out_c = boc[0]
    for i in range(len(boc)):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    mid("encoder.mid_block", boc[-1])
    gn("encoder.conv_norm_out", boc[-1])
    conv("encoder.conv_out", boc[-1], 2 * cfg.latent_channels, 3)
    conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    # decoder
    conv("decoder.conv_in", cfg.latent_channels, boc[-1], 3)
    mid("decoder.mid_block", boc[-1])
    rev = list(reversed(boc))
    out_c = rev[0]
    for i in range(len(boc)):
        prev, out_c = out_c, rev[i]
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out_c, out_c)
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    gn("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    return sh


def init_weights(cfg: VAEConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(cfg)
    w = {}
    for name, shape in shapes.items():
        is_norm = (".norm" in name) or ("group_norm" in name) or ("conv_norm_out" in name)
        if is_norm:
            w[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith(".weight") else 0.1 * torch.randn(shape, generator=g)
            continue
        wshape = shapes[name[: name.rfind(".")] + ".weight"]
        bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
        w[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return w


def _resnet(w, p, x, groups):
    h = F.silu(F.group_norm(x, groups, w[p + ".norm1.weight"], w[p + ".norm1.bias"], EPS))
    h = F.conv2d(h, w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, w[p + ".norm2.weight"], w[p + ".norm2.bias"], EPS))
    h = F.conv2d(h, w[p + ".conv2.weight"], w[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in w:
        x = F.conv2d(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
    return x + h


def _attn(w, p, x, groups):
    b, c, hh, ww = x.shape
    h = F.group_norm(x, groups, w[p + ".group_norm.weight"], w[p + ".group_norm.bias"], EPS)
    h = h.view(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, w[p + ".to_q.weight"], w[p + ".to_q.bias"])
    k = F.linear(h, w[p + ".to_k.weight"], w[p + ".to_k.bias"])
    v = F.linear(h, w[p + ".to_v.weight"], w[p + ".to_v.bias"])
    s = torch.softmax((q @ k.transpose(-1, -2)) * (c ** -0.5), dim=-1)   # one head of dim c
    o = F.linear(s @ v, w[p + ".to_out.0.weight"], w[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def _mid(w, p, x, groups):
    x = _resnet(w, p + ".resnets.0", x, groups)
    x = _attn(w, p + ".attentions.0", x, groups)
    return _resnet(w, p + ".resnets.1", x, groups)


def _tap(taps, name, h):
    if taps is not None:
        taps[name] = h.clone()
    return h


def _mid_t(w, p, x, groups, taps):
    x = _tap(taps, p + ".resnets.0", _resnet(w, p + ".resnets.0", x, groups))
    x = _tap(taps, p + ".attentions.0", _attn(w, p + ".attentions.0", x, groups))
    return _tap(taps, p + ".resnets.1", _resnet(w, p + ".resnets.1", x, groups))


def encode_moments(w, cfg: VAEConfig, x: torch.Tensor, taps=None) -> torch.Tensor:
    """quant_conv(encoder(x)): (B, 2*latent, H/8, W/8) = [mean | logvar]. `taps` (dict) collects per-block outputs."""
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = _tap(taps, "encoder.conv_in", F.conv2d(x, w["encoder.conv_in.weight"], w["encoder.conv_in.bias"], padding=1))
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = _tap(taps, f"encoder.down_blocks.{i}.resnets.{j}", _resnet(w, f"encoder.down_blocks.{i}.resnets.{j}", h, g))
        if i != len(boc) - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)     # Downsample2D(padding=0): asymmetric pad
            h = _tap(taps, p, F.conv2d(h, w[p + ".weight"], w[p + ".bias"], stride=2))
    h = _mid_t(w, "encoder.mid_block", h, g, taps)
    h = F.silu(F.group_norm(h, g, w["encoder.conv_norm_out.weight"], w["encoder.conv_norm_out.bias"], EPS))
    h = F.conv2d(h, w["encoder.conv_out.weight"], w["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, w["quant_conv.weight"], w["quant_conv.bias"])


def posterior_sample(moments: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """DiagonalGaussianDistribution(moments).sample(): mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean + torch.exp(0.5 * logvar) * noise


def decode(w, cfg: VAEConfig, z: torch.Tensor, taps=None) -> torch.Tensor:
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = F.conv2d(z, w["post_quant_conv.weight"], w["post_quant_conv.bias"])
    h = _tap(taps, "decoder.conv_in", F.conv2d(h, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], padding=1))
    h = _mid_t(w, "decoder.mid_block", h, g, taps)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block + 1):
            h = _tap(taps, f"decoder.up_blocks.{i}.resnets.{j}", _resnet(w, f"decoder.up_blocks.{i}.resnets.{j}", h, g))
        if i != len(boc) - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _tap(taps, p, F.conv2d(h, w[p + ".weight"], w[p + ".bias"], padding=1))
    h = F.silu(F.group_norm(h, g, w["decoder.conv_norm_out.weight"], w["decoder.conv_norm_out.bias"], EPS))
    return F.conv2d(h, w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], padding=1)
