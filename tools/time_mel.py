"""Mel codec timing on cuda:0 (CUDA events): encode 64 slices, decode 64 images (inverse mel + 32 Griffin-Lim iterations).
usage: python tools/time_mel.py [n] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_diffusion_b200.mel import Mel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
mel = Mel(x_res=256, y_res=256)
g = torch.Generator().manual_seed(0)
audio = (0.1 * torch.randn(n, mel.slice_size, generator=g)
         + 0.5 * torch.sin(2 * torch.pi * 440.0 * torch.arange(mel.slice_size) / 22050.0)[None]).to(dev)
imgs = mel.audio_slices_to_images(audio)
mel.images_to_audio(imgs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    imgs = mel.audio_slices_to_images(audio)
e1.record()
torch.cuda.synchronize()
enc = e0.elapsed_time(e1) / reps
e0.record()
for _ in range(reps):
    out = mel.images_to_audio(imgs)
e1.record()
torch.cuda.synchronize()
dec = e0.elapsed_time(e1) / reps
F, T, L = 1025, 256, mel.slice_size
enc_bytes = n * (L * 4 + 256 * 256)
dec_bytes = n * (32 * (3 * F * T * 16 + 2 * (T - 1) * 512 * 8) + F * T * 8 + 256 * 256)
print(json.dumps({"n": n, "encode_ms": enc, "decode_ms_incl_d2h": dec, "encode_gbs": enc_bytes / enc / 1e6,
                  "decode_gbs": dec_bytes / dec / 1e6, "decode_frac_of_6572.5": dec_bytes / dec / 1e6 / 6572.5}))
