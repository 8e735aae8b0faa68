"""Times AutoencoderKL encode / decode (config C4 shape: 256x256 <-> 32x32 latent) on cuda:0 with CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_diffusion_b200.vae import AutoencoderKL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
vae = AutoencoderKL(in_channels=1, out_channels=1, down_block_types=("DownEncoderBlock2D",) * 4,
                    up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                    layers_per_block=2, latent_channels=1, max_batch=B, seed=0).to(dev)
x = torch.randn(B, 1, 256, 256, device=dev).clamp(-1, 1)
z = torch.randn(B, 1, 32, 32, device=dev)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
te = timed(lambda: vae.encode(x).latent_dist.mode())
td = timed(lambda: vae.decode(z))
# encode 272 GF, decode 622 GF per sample
print(json.dumps({"batch": B, "encode_ms": te, "decode_ms": td, "encode_tflops": 272e-3 * B / te, "decode_tflops": 622e-3 * B / td,
                  "launches": vae.last_launch_count}))
