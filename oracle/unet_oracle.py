"""Plain-PyTorch fp32 restatement of diffusers==0.24.0 `UNet2DModel.forward`.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: diffusers is
not available in this environment; this follows the published 0.24.0 module
graph for exactly the configuration the reference constructs at
scripts/train_unet.py:115-137 and calls at
audiodiffusion/pipeline_audio_diffusion.py:163,237 and scripts/train_unet.py:257.

Weights are a flat dict keyed with the diffusers state-dict names
so a hub checkpoint would load unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Mirror of the UNet2DModel ctor kwargs used at scripts/train_unet.py:115-137."""

    sample_size: Tuple[int, int] = (256, 256)
    in_channels: int = 1
    out_channels: int = 1
    layers_per_block: int = 2
    block_out_channels: Tuple[int, ...] = (128, 128, 256, 256, 512, 512)
    down_block_types: Tuple[str, ...] = (
        "DownBlock2D", "DownBlock2D", "DownBlock2D", "DownBlock2D", "AttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = (
        "UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D")
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    attention_head_dim: int = 8
    # [3P-recall] remaining diffusers defaults: act_fn="silu", time_embedding_type="positional",
    # flip_sin_to_cos=True, freq_shift=0, downsample_padding=1, mid_block_scale_factor=1,
    # resnet_time_scale_shift="default", add_attention=True, dropout=0.


def param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """State-dict key -> shape, in the diffusers 0.24 layout."""
    sh: Dict[str, Tuple[int, ...]] = {}
    boc = cfg.block_out_channels
    temb = boc[0] * 4

    def conv(name, cin, cout, k):
        sh[name + ".weight"] = (cout, cin, k, k)
        sh[name + ".bias"] = (cout,)

    def lin(name, cin, cout):
        sh[name + ".weight"] = (cout, cin)
        sh[name + ".bias"] = (cout,)

    def gn(name, c):
        sh[name + ".weight"] = (c,)
        sh[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        gn(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", temb, cout)
        gn(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def attn(name, c):
        gn(name + ".group_norm", c)
        for p in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(name + "." + p, c, c)

    conv("conv_in", cfg.in_channels, boc[0], 3)
    lin("time_embedding.linear_1", boc[0], temb)
    lin("time_embedding.linear_2", temb, temb)
    out_c = boc[0]
    for i, typ in enumerate(cfg.down_block_types):
        in_c, out_c = out_c, boc[i]
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
            if typ == "AttnDownBlock2D":
                attn(f"down_blocks.{i}.attentions.{j}", out_c)
        if i != len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    resnet("mid_block.resnets.0", boc[-1], boc[-1])
    attn("mid_block.attentions.0", boc[-1])
    resnet("mid_block.resnets.1", boc[-1], boc[-1])
    rev = list(reversed(boc))
    out_c = rev[0]
    for i, typ in enumerate(cfg.up_block_types):
        prev_c, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        n = cfg.layers_per_block + 1
        for j in range(n):
            skip_c = in_c if j == n - 1 else out_c
            res_in = prev_c if j == 0 else out_c
            resnet(f"up_blocks.{i}.resnets.{j}", res_in + skip_c, out_c)
            if typ == "AttnUpBlock2D":
                attn(f"up_blocks.{i}.attentions.{j}", out_c)
        if i != len(boc) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    gn("conv_norm_out", boc[0])
    conv("conv_out", boc[0], cfg.out_channels, 3)
    return sh


def init_weights(cfg: UNetConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights, PyTorch default init (kaiming_uniform(a=sqrt 5) for conv/linear
    weights, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) biases, GroupNorm weight=1 bias=0), SURVEY §8(d).
    GroupNorm affine params are perturbed so that tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    shapes = param_shapes(cfg)
    for name, shape in shapes.items():
        is_norm = (".norm" in name) or ("group_norm" in name) or name.startswith("conv_norm_out")
        if is_norm:
            if name.endswith(".weight"):
                w[name] = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=dtype)
            else:
                w[name] = 0.1 * torch.randn(shape, generator=g, dtype=dtype)
            continue
        if name.endswith(".weight"):
            fan_in = int(math.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), ..)
            w[name] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
        else:
            wshape = shapes[name[:-5] + ".weight"]
            fan_in = int(math.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            w[name] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
    return w


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers `get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0, scale=1)`."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - 0.0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos
    return emb


def _resnet(w, p, x, temb_act, groups, eps, taps=None):
    h = F.group_norm(x, groups, w[p + ".norm1.weight"], w[p + ".norm1.bias"], eps)
    h = F.silu(h)
    if taps is not None:
        taps[p + ".act1"] = h
    h = F.conv2d(h, w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=1)
    t = F.linear(temb_act, w[p + ".time_emb_proj.weight"], w[p + ".time_emb_proj.bias"])
    h = h + t[:, :, None, None]
    if taps is not None:
        taps[p + ".h1"] = h
    h = F.group_norm(h, groups, w[p + ".norm2.weight"], w[p + ".norm2.bias"], eps)
    h = F.silu(h)
    h = F.conv2d(h, w[p + ".conv2.weight"], w[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in w:
        x = F.conv2d(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
    out = (x + h) / 1.0
    if taps is not None:
        taps[p] = out
    return out


def _attention(w, p, x, groups, eps, head_dim, taps=None):
    """diffusers `Attention` built from the deprecated AttnBlock config (residual_connection=True,
    upcast_softmax=True, rescale_output_factor=1) with AttnProcessor2_0 (SURVEY §8a U-attn)."""
    b, c, hh, ww = x.shape
    res = x
    h = F.group_norm(x, groups, w[p + ".group_norm.weight"], w[p + ".group_norm.bias"], eps)
    h = h.view(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, w[p + ".to_q.weight"], w[p + ".to_q.bias"])
    k = F.linear(h, w[p + ".to_k.weight"], w[p + ".to_k.bias"])
    v = F.linear(h, w[p + ".to_v.weight"], w[p + ".to_v.bias"])
    heads = c // head_dim
    q = q.view(b, -1, heads, head_dim).transpose(1, 2)
    k = k.view(b, -1, heads, head_dim).transpose(1, 2)
    v = v.view(b, -1, heads, head_dim).transpose(1, 2)
    s = torch.softmax((q @ k.transpose(-1, -2)) * (head_dim ** -0.5), dim=-1)
    o = (s @ v).transpose(1, 2).reshape(b, -1, c)
    o = F.linear(o, w[p + ".to_out.0.weight"], w[p + ".to_out.0.bias"])
    o = o.transpose(-1, -2).reshape(b, c, hh, ww)
    out = (o + res) / 1.0
    if taps is not None:
        taps[p] = out
    return out


def unet_forward(w: Dict[str, torch.Tensor], cfg: UNetConfig, sample: torch.Tensor, timestep,
                 taps: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """UNet2DModel.forward(sample, timestep).sample  (pipeline_audio_diffusion.py:163).

    `timestep`: python int, 0-d tensor, or 1-D LongTensor (per-sample, train_unet.py:257).
    `taps`: optional dict filled with named intermediate activations (for per-layer parity tests).
    """
    boc = cfg.block_out_channels
    g, eps = cfg.norm_num_groups, cfg.norm_eps
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.long, device=sample.device)
    elif t.ndim == 0:
        t = t[None].to(sample.device)
    t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
    emb = timestep_embedding(t, boc[0]).to(sample.dtype)
    emb = F.linear(emb, w["time_embedding.linear_1.weight"], w["time_embedding.linear_1.bias"])
    emb = F.silu(emb)
    emb = F.linear(emb, w["time_embedding.linear_2.weight"], w["time_embedding.linear_2.bias"])
    temb_act = F.silu(emb)  # every ResnetBlock2D applies nonlinearity(temb) before time_emb_proj
    if taps is not None:
        taps["temb_act"] = temb_act

    h = F.conv2d(sample, w["conv_in.weight"], w["conv_in.bias"], padding=1)
    if taps is not None:
        taps["conv_in"] = h
    skips: List[torch.Tensor] = [h]
    for i, typ in enumerate(cfg.down_block_types):
        for j in range(cfg.layers_per_block):
            h = _resnet(w, f"down_blocks.{i}.resnets.{j}", h, temb_act, g, eps, taps)
            if typ == "AttnDownBlock2D":
                h = _attention(w, f"down_blocks.{i}.attentions.{j}", h, g, eps, cfg.attention_head_dim, taps)
            skips.append(h)
        if i != len(boc) - 1:
            p = f"down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(h, w[p + ".weight"], w[p + ".bias"], stride=2, padding=1)
            if taps is not None:
                taps[p] = h
            skips.append(h)
    h = _resnet(w, "mid_block.resnets.0", h, temb_act, g, eps, taps)
    h = _attention(w, "mid_block.attentions.0", h, g, eps, cfg.attention_head_dim, taps)
    h = _resnet(w, "mid_block.resnets.1", h, temb_act, g, eps, taps)
    for i, typ in enumerate(cfg.up_block_types):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = _resnet(w, f"up_blocks.{i}.resnets.{j}", h, temb_act, g, eps, taps)
            if typ == "AttnUpBlock2D":
                h = _attention(w, f"up_blocks.{i}.attentions.{j}", h, g, eps, cfg.attention_head_dim, taps)
        if i != len(boc) - 1:
            p = f"up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, w[p + ".weight"], w[p + ".bias"], padding=1)
            if taps is not None:
                taps[p] = h
    assert not skips
    h = F.group_norm(h, g, w["conv_norm_out.weight"], w["conv_norm_out.bias"], eps)
    h = F.silu(h)
    h = F.conv2d(h, w["conv_out.weight"], w["conv_out.bias"], padding=1)
    return h


def unet_flops(cfg: UNetConfig, h: int, w: int) -> float:
    """Analytic 2*MAC count per sample per forward (convs + linears + attention); SURVEY §8(d)."""
    total = 0.0
    shapes = param_shapes(cfg)
    boc = cfg.block_out_channels
    # resolution of every conv: replay the graph
    res = {}
    hh, ww = h, w
    res["conv_in"] = (hh, ww)
    for i, typ in enumerate(cfg.down_block_types):
        for j in range(cfg.layers_per_block):
            res[f"down_blocks.{i}.resnets.{j}"] = (hh, ww)
            res[f"down_blocks.{i}.attentions.{j}"] = (hh, ww)
        if i != len(boc) - 1:
            hh, ww = hh // 2, ww // 2
            res[f"down_blocks.{i}.downsamplers.0.conv"] = (hh, ww)
    res["mid_block"] = (hh, ww)
    for i, typ in enumerate(cfg.up_block_types):
        for j in range(cfg.layers_per_block + 1):
            res[f"up_blocks.{i}.resnets.{j}"] = (hh, ww)
            res[f"up_blocks.{i}.attentions.{j}"] = (hh, ww)
        if i != len(boc) - 1:
            hh, ww = hh * 2, ww * 2
            res[f"up_blocks.{i}.upsamplers.0.conv"] = (hh, ww)
    res["conv_out"] = (hh, ww)
    for name, shape in shapes.items():
        if not name.endswith(".weight"):
            continue
        key = None
        for k in res:
            if name.startswith(k):
                key = k if key is None or len(k) > len(key) else key
        if len(shape) == 4:
            rh, rw = res[key]
            total += 2.0 * shape[0] * shape[1] * shape[2] * shape[3] * rh * rw
        elif len(shape) == 2:
            if "time_emb" in name:
                total += 2.0 * shape[0] * shape[1]
            else:  # attention linears: per token
                rh, rw = res[key]
                total += 2.0 * shape[0] * shape[1] * rh * rw
    for k, (rh, rw) in res.items():
        if "attentions" in k and (k + ".to_q.weight") in shapes:
            c = shapes[k + ".to_q.weight"][0]
            seq = rh * rw
            total += 2.0 * 2.0 * seq * seq * c  # QK^T and PV
    return total
