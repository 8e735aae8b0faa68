"""Times one training iteration (scripts/train_unet.py:238-267: add_noise, U-Net forward + backward, clip + AdamW + EMA) of
the reference U-Net at 256x256 on cuda:0 with CUDA events.  usage: python tools/train_bench.py [batch] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_diffusion_b200.schedulers import DDPMScheduler
from audio_diffusion_b200.training import EMAModel, FusedAdamW, train_step
from audio_diffusion_b200.unet import UNet2DModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = torch.device("cuda:0")
model = UNet2DModel(sample_size=(HW, HW), in_channels=1, out_channels=1, layers_per_block=2,
                    block_out_channels=(128, 128, 256, 256, 512, 512),
                    down_block_types=("DownBlock2D",) * 4 + ("AttnDownBlock2D", "DownBlock2D"),
                    up_block_types=("UpBlock2D", "AttnUpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D", "UpBlock2D"),
                    seed=0).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-4, betas=(0.95, 0.999), weight_decay=1e-6, eps=1e-8, max_grad_norm=1.0)
ema = EMAModel(model.parameters(), inv_gamma=1.0, power=0.75, max_value=0.9999)
opt.attach_ema(ema)
sch = DDPMScheduler()
g = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 1, HW, HW, device=dev, generator=g) * 2 - 1
losses = []
def step():
    losses.append(train_step(model, opt, sch, x, ema=ema, generator=g))
for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(STEPS):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / STEPS
gf = {256: 496.42, 64: 31.0, 32: 7.76}.get(HW)
print(json.dumps({"batch": B, "res": HW, "ms_per_train_step": ms, "images_per_s": B / ms * 1e3,
                  "tflops_3x_forward": (3 * gf * B / ms) if gf else None,
                  "backward_launches": model.last_backward_launch_count, "forward_launches": model.last_launch_count,
                  "loss": [float(l) for l in losses], "max_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
