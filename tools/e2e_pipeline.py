"""End-to-end timing of AudioDiffusionPipeline.__call__ (noise -> uint8 images -> audio) on one GPU.

    python tools/e2e_pipeline.py --scheduler ddim --steps 50 --batch 64      # BASELINE config C3, per-GPU share
    python tools/e2e_pipeline.py --scheduler ddpm --steps 40 --batch 64      # 40 of the 1000 DDPM steps + full tail

Reports wall time of the whole call and of its stages (denoise loop, float->uint8 + PIL, batched Griffin-Lim)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_diffusion_b200.mel import Mel
from audio_diffusion_b200.pipeline import AudioDiffusionPipeline
from audio_diffusion_b200.schedulers import DDIMScheduler, DDPMScheduler
from audio_diffusion_b200.unet import UNet2DModel
from bench import REF_ARCH

ap = argparse.ArgumentParser()
ap.add_argument("--scheduler", default="ddim")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--res", type=int, default=256)
args = ap.parse_args()
dev = torch.device("cuda:0")
unet = UNet2DModel(sample_size=(args.res, args.res), seed=0, **REF_ARCH).to(dev)
sch = DDIMScheduler() if args.scheduler == "ddim" else DDPMScheduler()
mel = Mel(x_res=args.res, y_res=args.res)
pipe = AudioDiffusionPipeline(vqvae=None, unet=unet, mel=mel, scheduler=sch)
pipe.set_progress_bar_config(disable=True)
gen = torch.Generator(device=dev).manual_seed(42)
pipe(batch_size=args.batch, steps=2, generator=gen)  # warm-up (binds workspace, packs weights)
torch.cuda.synchronize()

t0 = time.perf_counter()
imgs = pipe(batch_size=args.batch, steps=args.steps, generator=gen, return_audio=False)
torch.cuda.synchronize()
t_img = time.perf_counter() - t0
u8 = torch.stack([torch.from_numpy(__import__("numpy").asarray(i)) for i in imgs]).to(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
audio = mel.images_to_audio(u8)
torch.cuda.synchronize()
t_audio = time.perf_counter() - t0
t0 = time.perf_counter()
out = pipe(batch_size=args.batch, steps=args.steps, generator=gen)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
full_steps = 50 if args.scheduler == "ddim" else 1000
print(json.dumps({
    "scheduler": args.scheduler, "steps_run": args.steps, "batch": args.batch, "res": args.res,
    "call_s": t_all, "images_only_s": t_img, "griffin_lim_batch_s": t_audio,
    "audio_shape": list(out.audios.shape),
    "ms_per_step": 1e3 * t_img / args.steps,
    "projected_full_call_s": t_img * full_steps / args.steps + t_audio,
    "projected_mel_spectrograms_per_s": args.batch / (t_img * full_steps / args.steps + t_audio)}))
