"""`AutoencoderKL` — drop-in for `diffusers.AutoencoderKL` on the two calls the reference makes
(audiodiffusion/pipeline_audio_diffusion.py:143-147 `vqvae.encode(x).latent_dist.sample(generator=...)`,
:187-190 `vqvae.decode(z)["sample"]`; scripts/train_unet.py:99-104, :230-235), in the architecture of
config/ldm_autoencoder_kl.yaml:18-28 and with the state-dict keys audiodiffusion/utils.py:156-303
(`convert_ldm_to_hf_vae`) produces.

Encoder, decoder, quant/post-quant convs and the posterior sampling run in libb200ad.so (vae.cu); PyTorch owns the
parameters (fp32 `nn.Parameter`s), the packed bf16 weights and the activation workspace.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import MAX_BLOCKS, VAEConfigC
from .unet import _Cfg, _set_deep


class DecoderOutput(dict):
    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class DiagonalGaussianDistribution:
    """Posterior returned by `encode(x).latent_dist`: `.sample(generator)`, `.mode()`, `.mean`, `.logvar`, `.std`, `.var`.

    `sample()` draws its noise exactly as diffusers does (`randn_tensor(mean.shape, generator, device)`), then
    mean + std * noise is evaluated by the encoder's tail kernel (vae_sample_kernel) — the moments never leave the device.
    """

    def __init__(self, vae: "AutoencoderKL", x: torch.Tensor):
        self._vae = vae
        self._x = x
        self._moments: Optional[torch.Tensor] = None

    def _run(self, noise: Optional[torch.Tensor]) -> torch.Tensor:
        z, m = self._vae._encode(self._x, noise)
        self._moments = m
        return z

    @property
    def parameters(self) -> torch.Tensor:
        if self._moments is None:
            self._run(None)
        return self._moments

    @property
    def mean(self) -> torch.Tensor:
        return torch.chunk(self.parameters, 2, dim=1)[0]

    @property
    def logvar(self) -> torch.Tensor:
        return torch.clamp(torch.chunk(self.parameters, 2, dim=1)[1], -30.0, 20.0)

    @property
    def std(self) -> torch.Tensor:
        return torch.exp(0.5 * self.logvar)

    @property
    def var(self) -> torch.Tensor:
        return torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        x = self._x
        shape = self._vae.latent_shape(x.shape)
        gdev = generator.device if generator is not None else x.device
        noise = torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to(x.device)
        return self._run(noise)

    def mode(self) -> torch.Tensor:
        return self._run(None)


class AutoencoderKLOutput(dict):
    def __init__(self, latent_dist):
        super().__init__(latent_dist=latent_dist)
        self.latent_dist = latent_dist


class AutoencoderKL(nn.Module):
    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        down_block_types: Sequence[str] = ("DownEncoderBlock2D",),
        up_block_types: Sequence[str] = ("UpDecoderBlock2D",),
        block_out_channels: Sequence[int] = (64,),
        layers_per_block: int = 1,
        act_fn: str = "silu",
        latent_channels: int = 4,
        norm_num_groups: int = 32,
        sample_size: int = 32,
        scaling_factor: float = 0.18215,
        max_batch: int = 16,
        seed: Optional[int] = None,
    ):
        super().__init__()
        bad = []
        if act_fn != "silu": bad.append("act_fn")
        if any(t != "DownEncoderBlock2D" for t in down_block_types): bad.append("down_block_types")
        if any(t != "UpDecoderBlock2D" for t in up_block_types): bad.append("up_block_types")
        if len(block_out_channels) > MAX_BLOCKS or len(down_block_types) != len(block_out_channels): bad.append("blocks")
        if bad:
            raise ValueError(f"AutoencoderKL(b200): unsupported configuration: {bad}")
        self.config = _Cfg(
            in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, act_fn=act_fn, latent_channels=latent_channels,
            norm_num_groups=norm_num_groups, sample_size=sample_size, scaling_factor=scaling_factor,
            _class_name="AutoencoderKL")
        self.max_batch = int(max_batch)  # activations are bound for at most this many images; larger batches are chunked
        c = VAEConfigC()
        c.in_channels, c.out_channels, c.latent_channels = in_channels, out_channels, latent_channels
        c.layers_per_block, c.num_blocks = layers_per_block, len(block_out_channels)
        for i, v in enumerate(block_out_channels):
            c.block_out_channels[i] = int(v)
        c.norm_num_groups, c.norm_eps = norm_num_groups, 1e-6
        self._c = c
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.b200ad_vae_create(C.byref(c), C.byref(h)))
        self._h = h
        self._factor = 1 << (len(block_out_channels) - 1)
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        self._pnames = []
        dims = (C.c_int64 * 4)()
        shapes: Dict[str, Tuple[int, ...]] = {}
        for i in range(L.b200ad_vae_num_params(h)):
            name = L.b200ad_vae_param_name(h, i).decode()
            nd = L.b200ad_vae_param_shape(h, i, dims)
            shapes[name] = tuple(int(dims[k]) for k in range(nd))
            self._pnames.append(name)
        for name in self._pnames:
            shape = shapes[name]
            is_norm = (".norm" in name) or ("group_norm" in name) or ("conv_norm_out" in name)
            if is_norm:
                t = torch.ones(shape) if name.endswith(".weight") else torch.zeros(shape)
            else:
                wshape = shapes[name[: name.rfind(".")] + ".weight"]
                bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            _set_deep(self, name, nn.Parameter(t))
        self._packed = None
        self._packed_key = None
        self._ws = None
        self._ws_key = None

    # ------------------------------------------------------------------ diffusers directory layout
    _KEEP = ("in_channels", "out_channels", "down_block_types", "up_block_types", "block_out_channels",
             "layers_per_block", "act_fn", "latent_channels", "norm_num_groups", "sample_size", "scaling_factor")

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw) -> "AutoencoderKL":
        """`AutoencoderKL.from_pretrained(dir)` (scripts/train_unet.py:99-104): config.json + diffusion_pytorch_model.*;
        raises EnvironmentError when the directory holds no model, which the reference catches to fall back to the
        pipeline's `vqvae` component."""
        import json
        import os
        sub = os.path.join(path, subfolder) if subfolder else path
        cfgp = os.path.join(sub, "config.json")
        if not os.path.exists(cfgp):
            raise EnvironmentError(f"{sub} does not contain an AutoencoderKL (config.json missing)")
        with open(cfgp) as f:
            cfg = json.load(f)
        model = cls(**{k: cfg[k] for k in cls._KEEP if k in cfg})
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
            if v.dim() == 4 and k.endswith(".weight") and (".to_" in k) and v.shape[2:] == (1, 1):
                v = v[:, :, 0, 0]  # ldm-converted attention projections are 1x1 convs (utils.py:285-303)
            fixed[k] = v.to(torch.float32)
        model.load_state_dict(fixed)
        return model

    _DEPRECATED = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the deprecated attention key names (`query/key/value/proj_attn`, possibly conv-shaped) that
        audiodiffusion/utils.py:33-60,120-129 still writes — diffusers converts them on load
        ([3P-recall] `_convert_deprecated_attention_blocks`)."""
        fixed = {}
        for k, v in state_dict.items():
            for a, b in self._DEPRECATED.items():
                k = k.replace(a, b)
            if ".attentions." in k and k.endswith(".weight") and v.dim() > 2:
                v = v.reshape(v.shape[0], v.shape[1])
            fixed[k] = v
        return super().load_state_dict(fixed, strict=strict, **kw)

    def save_pretrained(self, path: str) -> None:
        import json
        import os
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump({k: v for k, v in self.config.items()}, f, indent=2)
        from safetensors.torch import save_file
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(path, "diffusion_pytorch_model.safetensors"))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().b200ad_vae_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def latent_shape(self, image_shape) -> Tuple[int, int, int, int]:
        n, _, hh, ww = image_shape
        return (n, self.config.latent_channels, hh // self._factor, ww // self._factor)

    # ------------------------------------------------------------------ engine plumbing
    def _ensure_bound(self, n: int, hh: int, ww: int) -> None:
        _lib.require_cuda()
        L = _lib.lib()
        named = dict(self.named_parameters())
        params = [named[k] for k in self._pnames]
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.B200ADError("AutoencoderKL(b200): parameters must live on a CUDA device (call .to('cuda'))")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.B200ADError("AutoencoderKL(b200): parameters must be contiguous fp32")
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed.device != dev:
            self._packed = torch.empty(L.b200ad_vae_packed_bytes(self._h), dtype=torch.uint8, device=dev)
            self._packed_key = None
            self._ws_key = None
        if key != self._packed_key:
            arr = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            _lib.check(L.b200ad_vae_set_params(self._h, arr, self._packed.data_ptr(), self._packed.numel(),
                                               _lib.stream_ptr()))
            self._packed_key = key
            self._ws_key = None
        wkey = (n, hh, ww, dev)
        if wkey != self._ws_key:
            need = L.b200ad_vae_workspace_bytes(self._h, n, hh, ww)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(L.b200ad_vae_bind_workspace(self._h, self._ws.data_ptr(), self._ws.numel(), n, hh, ww,
                                                   _lib.stream_ptr()))
            self._ws_key = wkey

    def _check(self, x: torch.Tensor, channels: int, what: str) -> torch.Tensor:
        _lib.require_cuda()
        if x.device.type != "cuda":
            raise _lib.B200ADError(f"AutoencoderKL(b200): {what} must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != channels:
            raise ValueError(f"AutoencoderKL(b200): {what} must be (N, {channels}, H, W), got {tuple(x.shape)}")
        return x.to(torch.float32).contiguous()

    @torch.no_grad()
    def _encode(self, x: torch.Tensor, noise: Optional[torch.Tensor]):
        x = self._check(x, self.config.in_channels, "encode input")
        n, _, hh, ww = x.shape
        if hh % self._factor or ww % self._factor:
            raise ValueError(f"AutoencoderKL(b200): H and W must be multiples of {self._factor}")
        lshape = self.latent_shape(x.shape)
        z = torch.empty(lshape, dtype=torch.float32, device=x.device)
        m = torch.empty((n, 2 * lshape[1], lshape[2], lshape[3]), dtype=torch.float32, device=x.device)
        if noise is not None:
            noise = noise.to(device=x.device, dtype=torch.float32).contiguous()
            if tuple(noise.shape) != tuple(lshape):
                raise ValueError("noise must have the latent shape")
        L = _lib.lib()
        with torch.cuda.device(x.device):
            for s in range(0, n, self.max_batch):
                e = min(n, s + self.max_batch)
                self._ensure_bound(e - s, hh, ww)
                _lib.check(L.b200ad_vae_encode(self._h, x[s:e].data_ptr(),
                                               noise[s:e].data_ptr() if noise is not None else None,
                                               z[s:e].data_ptr(), m[s:e].data_ptr(), _lib.stream_ptr()))
        return z, m

    # ------------------------------------------------------------------ public calls
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """`vqvae.encode(x).latent_dist` — the encoder runs when the distribution is sampled / inspected."""
        dist = DiagonalGaussianDistribution(self, self._check(x, self.config.in_channels, "encode input"))
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(dist)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """`vqvae.decode(z)["sample"]` (pipeline_audio_diffusion.py:190)."""
        z = self._check(z, self.config.latent_channels, "latents")
        n, _, lh, lw = z.shape
        hh, ww = lh * self._factor, lw * self._factor
        out = torch.empty((n, self.config.out_channels, hh, ww), dtype=torch.float32, device=z.device)
        L = _lib.lib()
        with torch.cuda.device(z.device):
            for s in range(0, n, self.max_batch):
                e = min(n, s + self.max_batch)
                self._ensure_bound(e - s, hh, ww)
                _lib.check(L.b200ad_vae_decode(self._h, z[s:e].data_ptr(), out[s:e].data_ptr(), _lib.stream_ptr()))
        if not return_dict:
            return (out,)
        return DecoderOutput(out)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator=generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)

    def debug_tensor(self, name: str) -> torch.Tensor:
        L = _lib.lib()
        dims = (C.c_int * 3)()
        _lib.check(min(0, L.b200ad_vae_debug_tensor(self._h, name.encode(), None, dims, _lib.stream_ptr())))
        n = self._ws_key[0]
        out = torch.empty((n, dims[0], dims[1], dims[2]), dtype=torch.float32, device=self.device)
        _lib.check(min(0, L.b200ad_vae_debug_tensor(self._h, name.encode(), out.data_ptr(), dims, _lib.stream_ptr())))
        return out

    @property
    def last_launch_count(self) -> int:
        return _lib.lib().b200ad_vae_last_launch_count(self._h)
