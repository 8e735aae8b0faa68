"""Time one conv_tc launch configuration (CUDA events) — used with B200AD_CONV_DBG experiments and under ncu."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_diffusion_b200 import _lib

N, cin, cout, H, W, K = [int(a) for a in sys.argv[1:7]] if len(sys.argv) > 6 else (64, 128, 128, 256, 256, 3)
reps = int(os.environ.get("REPS", "5"))
res = os.environ.get("RES", "0") == "1"
L = _lib.lib()
dev = torch.device("cuda:0")
x = torch.randn(N, cin, H, W, device=dev)
w = torch.randn(cout, cin, K, K, device=dev) * 0.05
b = torch.randn(cout, device=dev)
y = torch.empty(N, cout, H, W, device=dev)
r = torch.randn(N, cout, H, W, device=dev) if res else None
stats = torch.empty(N, cout // 4, 2, device=dev)
nb = L.b200ad_conv2d_scratch_bytes(N, cin, cout, H, W, K, 1)
scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr()


def run():
    _lib.check(L.b200ad_conv2d(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, r.data_ptr() if res else None,
                               y.data_ptr(), stats.data_ptr(), N, cin, cout, H, W, K, 1, scratch.data_ptr(), nb, st))


run()
torch.cuda.synchronize()
# the op-level entry runs layout conversions around the conv; time the whole thing and the conversions separately is
# not possible from here, so rely on ncu / nsys-free event timing of the conv kernel via the profile API in bench.py.
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
print(f"dbg={os.environ.get('B200AD_CONV_DBG', '0')} N={N} cin={cin} cout={cout} {H}x{W} K={K} res={res}: "
      f"{e0.elapsed_time(e1) / reps:.3f} ms per op-level call (includes layout conversions)")
