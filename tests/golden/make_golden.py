"""Generates tests/golden/* by running the reference's OWN Python files (read from /root/reference, unchanged) in this
container.  Run from the repo root:  python tests/golden/make_golden.py

 * vae_key_map.json — output of audiodiffusion/utils.py::convert_ldm_vae_checkpoint on an ldm-format checkpoint of the
   config/ldm_autoencoder_kl.yaml shape: [hf key, shape] in the order the reference writes them.  Pins the parameter
   table of libb200ad's AutoencoderKL (b200ad_vae_param_name / _shape).
 * pipeline_ddpm_small.npz — audiodiffusion/pipeline_audio_diffusion.py::AudioDiffusionPipeline.__call__ (the reference
   file, on the import shim) driving the fp32 CPU oracle U-Net (seeded synthetic weights, oracle/unet_oracle.py) and the
   ORACLE's DDPM scheduler (oracle/schedulers_oracle.py - nothing of the product computes a number in this fixture; the
   product's own scheduler gives bit-identical images, which the generator asserts) for 6 steps from a given noise tensor
   with a CPU step generator: the uint8 images it returns.
   The GPU pipeline must reproduce them within the stated tolerance (tests/test_gpu_pipeline.py).
The numerical content is the oracle's (diffusers cannot be imported here); what the fixtures pin is the reference
files' own logic: key mapping, loop order, noise draws, float->uint8 conversion.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio_diffusion_b200", "compat"), REF, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

SMALL = dict(in_channels=1, out_channels=1, layers_per_block=2, block_out_channels=(128, 256),
             down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))


def vae_key_map():
    from audiodiffusion.utils import convert_ldm_vae_checkpoint
    from oracle.vae_oracle import VAEConfig, init_weights
    from test_cpu_dropin import _hf_to_ldm_vae
    w = init_weights(VAEConfig(), seed=0)
    conv = convert_ldm_vae_checkpoint(_hf_to_ldm_vae(w), None)
    out = [[k, list(v.shape)] for k, v in conv.items()]
    with open(os.path.join(ROOT, "tests", "golden", "vae_key_map.json"), "w") as f:
        json.dump({"source": "audiodiffusion/utils.py:156-291 convert_ldm_vae_checkpoint", "keys": out}, f, indent=0)
    print("vae_key_map.json", len(out))


def pipeline_small():
    from audiodiffusion.pipeline_audio_diffusion import AudioDiffusionPipeline as RefPipe
    from diffusers import DDPMScheduler as ProductDDPM      # the import shim = the product's scheduler (cross-check only)
    from oracle.schedulers_oracle import OracleDDPM
    from oracle.unet_oracle import UNetConfig, init_weights, unet_forward

    class OracleUNet:
        def __init__(self):
            self.cfg = UNetConfig(sample_size=(32, 32), **SMALL)
            self.w = init_weights(self.cfg, seed=11)
            self.sample_size = (32, 32)
            self.in_channels = 1

        def __call__(self, x, t):
            return {"sample": unet_forward(self.w, self.cfg, x, t)}

    class FakeMel:
        x_res, y_res, hop_length = 32, 32, 512

        def get_sample_rate(self):
            return 22050

        def image_to_audio(self, image):
            return np.zeros((self.x_res - 1) * self.hop_length, dtype=np.float32)

    noise = torch.randn(2, 1, 32, 32, generator=torch.Generator().manual_seed(42))

    def run(scheduler):
        pipe = RefPipe(vqvae=None, unet=OracleUNet(), mel=FakeMel(), scheduler=scheduler)
        pipe.set_progress_bar_config(disable=True)
        images, _ = pipe(batch_size=2, steps=6, noise=noise.clone(), step_generator=torch.Generator().manual_seed(7),
                         return_dict=False)
        return np.stack([np.asarray(im) for im in images]).astype(np.uint8)
    arr = run(OracleDDPM())
    assert np.array_equal(arr, run(ProductDDPM())), "the product's DDPMScheduler disagrees with the oracle's"

    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pipeline_ddpm_small.npz"), noise=noise.numpy(), images=arr,
                        steps=6, weight_seed=11, step_seed=7)
    print("pipeline_ddpm_small.npz", arr.shape, arr.mean())


if __name__ == "__main__":
    vae_key_map()
    pipeline_small()
