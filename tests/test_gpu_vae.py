"""GPU parity of the latent autoencoder (C ABI `b200ad_vae_encode` / `b200ad_vae_decode`) against oracle/vae_oracle.py.

Tolerance (stated): bf16 activations / GEMM operands with fp32 accumulation against an fp32 oracle.
 * encoder (blocks, moments, latents): max|err| <= 6% of the oracle tensor's max-abs, rms error <= 1.5% of its rms — the
   U-Net bar;
 * decoder: a chain of ~30 convolutions with no skip connections, so bf16 rounding compounds: the fp32 oracle with its
   conv operands/outputs merely ROUNDED to bf16 already differs from itself by 2.1-3.4% rms at the decoder output
   (tests/test_cpu_oracle.py::test_vae_bf16_rounding_floor measures it).  Bar: rms <= 3% per decoder block, <= 5% for the
   decoded image, max|err| <= 10%.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(cuda, seed=0, **kw):
    from audio_diffusion_b200.vae import AutoencoderKL
    from oracle.vae_oracle import VAEConfig, init_weights
    ocfg = VAEConfig(**kw)
    w = init_weights(ocfg, seed=seed)
    n = len(ocfg.block_out_channels)
    model = AutoencoderKL(in_channels=ocfg.in_channels, out_channels=ocfg.out_channels,
                          down_block_types=("DownEncoderBlock2D",) * n, up_block_types=("UpDecoderBlock2D",) * n,
                          block_out_channels=ocfg.block_out_channels, layers_per_block=ocfg.layers_per_block,
                          latent_channels=ocfg.latent_channels, norm_num_groups=ocfg.norm_num_groups)
    assert set(model.state_dict().keys()) == set(w.keys())
    model.load_state_dict(w)
    return model.to(cuda), ocfg, w


def _cmp(name, got, ref, max_tol=6e-2, rms_tol=1.5e-2):
    err = got - ref
    mx = err.abs().max().item() / (ref.abs().max().item() + 1e-12)
    rms = (err.pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-12)).item()
    assert mx <= max_tol and rms <= rms_tol, f"{name}: max-rel {mx:.4f} rms-rel {rms:.4f}"
    return mx, rms


DEC_BLOCK = dict(max_tol=1e-1, rms_tol=3e-2)
DEC_OUT = dict(max_tol=1e-1, rms_tol=5e-2)


def test_vae_layers_small(cuda):
    """Every block of encoder and decoder at 64x64 (latents 8x8), buffers un-pooled so each tap survives."""
    from oracle.vae_oracle import decode, encode_moments, posterior_sample
    os.environ["B200AD_DEBUG_NOPOOL"] = "1"
    try:
        model, ocfg, w = _build(cuda)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(2, 1, 64, 64, generator=g).clamp(-1, 1)
        taps = {}
        m_ref = encode_moments(w, ocfg, x, taps)
        post = model.encode(x.to(cuda)).latent_dist
        gz = torch.Generator().manual_seed(11)
        z = post.sample(generator=gz)
        noise = torch.randn(z.shape, generator=torch.Generator().manual_seed(11))
        report = []
        for name in taps:
            report.append((name,) + _cmp(name, model.debug_tensor(name).cpu(), taps[name]))
        report.append(("moments",) + _cmp("moments", post.parameters.cpu(), m_ref))
        z_ref = posterior_sample(m_ref, noise)
        report.append(("z",) + _cmp("z", z.cpu(), z_ref))
        # decoder from the ORACLE latents (teacher-forced), so decoder error is not compounded with the encoder's
        dtaps = {}
        y_ref = decode(w, ocfg, z_ref, dtaps)
        y = model.decode(z_ref.to(cuda))["sample"]
        for r in report:
            print("%-48s max-rel %.4f rms-rel %.4f" % r)
        for name in dtaps:
            r = (name,) + _cmp(name, model.debug_tensor(name).cpu(), dtaps[name], **DEC_BLOCK)
            print("%-48s max-rel %.4f rms-rel %.4f" % r)
        print("%-48s max-rel %.4f rms-rel %.4f" % (("decode",) + _cmp("decode", y.cpu(), y_ref, **DEC_OUT)))
    finally:
        os.environ.pop("B200AD_DEBUG_NOPOOL", None)


def test_vae_pooled_nonsquare_and_mode(cuda):
    """Pooled buffers, non-square input, batch chunking (max_batch) and `.mode()` == mean of the moments."""
    from oracle.vae_oracle import decode, encode_moments
    model, ocfg, w = _build(cuda, seed=2)
    model.max_batch = 2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 1, 32, 96, generator=g).clamp(-1, 1)
    m_ref = encode_moments(w, ocfg, x)
    post = model.encode(x.to(cuda)).latent_dist
    zm = post.mode()
    _cmp("mode", zm.cpu(), m_ref[:, :1])
    _cmp("moments", post.parameters.cpu(), m_ref)
    y = model.decode(m_ref[:, :1].to(cuda).contiguous())["sample"]
    _cmp("decode", y.cpu(), decode(w, ocfg, m_ref[:, :1]), **DEC_OUT)
    assert y.shape == (3, 1, 32, 96)
    assert model.last_launch_count > 0


@pytest.mark.timeout(600)
def test_vae_c4_resolution(cuda):
    """config C4: 256x256 mel image <-> 32x32 latent, batch 1 (the oracle needs ~10 s on CPU)."""
    from oracle.vae_oracle import decode, encode_moments
    model, ocfg, w = _build(cuda, seed=4)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 1, 256, 256, generator=g).clamp(-1, 1)
    m_ref = encode_moments(w, ocfg, x)
    post = model.encode(x.to(cuda)).latent_dist
    zm = post.mode()
    assert zm.shape == (1, 1, 32, 32)
    _cmp("moments", post.parameters.cpu(), m_ref)
    z = m_ref[:, :1].contiguous()
    y = model.decode(z.to(cuda))["sample"]
    _cmp("decode", y.cpu(), decode(w, ocfg, z), **DEC_OUT)


def test_vae_rgb_hub_shape(cuda):
    """Hub VAEs are 3-channel RGB with 4 latent channels (scripts/train_unet.py:81-82, pipeline_audio_diffusion.py:198):
    the generic conv_in (cin = 3 / 4), the 8-wide moments tail and the cout = 3 output kernel."""
    from oracle.vae_oracle import decode, encode_moments, posterior_sample
    model, ocfg, w = _build(cuda, seed=9, in_channels=3, out_channels=3, latent_channels=4)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 32, 32, generator=g).clamp(-1, 1)
    m_ref = encode_moments(w, ocfg, x)
    post = model.encode(x.to(cuda)).latent_dist
    z = post.sample(generator=torch.Generator().manual_seed(4))
    noise = torch.randn(z.shape, generator=torch.Generator().manual_seed(4))
    _cmp("moments", post.parameters.cpu(), m_ref)
    _cmp("z", z.cpu(), posterior_sample(m_ref, noise), max_tol=6e-2, rms_tol=1.5e-2)
    zr = m_ref[:, :4].contiguous()
    y = model.decode(zr.to(cuda))["sample"]
    assert y.shape == (2, 3, 32, 32)
    _cmp("decode", y.cpu(), decode(w, ocfg, zr), **DEC_OUT)
