"""Structured probe of the tcgen05 wgrad kernel's operand mapping (MN-major descriptors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_diffusion_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
N, cin, cout, H, W, K = 1, 32, 128, 8, 8, 1


def run(gy, a, variant):
    os.environ["B200AD_WGRAD_VARIANT"] = str(variant)
    dw = torch.full((cout, cin, K, K), float("nan"), device=dev)
    nb = L.b200ad_conv2d_wgrad_scratch_bytes(N, cin, cout, H, W)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    gd, ad = gy.to(dev).contiguous(), a.to(dev).contiguous()
    _lib.check(L.b200ad_conv2d_wgrad(gd.data_ptr(), ad.data_ptr(), dw.data_ptr(), N, cin, cout, H, W, K,
                                     scratch.data_ptr(), nb, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return dw.cpu()[:, :, 0, 0]


for variant in (0, 1):
    print("=== variant", variant)
    # A: channel mapping. one pixel (h=2,w=3) carries (co+1) in gy and (ci+1) in a  -> expect (co+1)*(ci+1)
    gy = torch.zeros(N, cout, H, W); a = torch.zeros(N, cin, H, W)
    gy[0, :, 2, 3] = torch.arange(1, cout + 1).float()
    a[0, :, 2, 3] = torch.arange(1, cin + 1).float()
    d = run(gy, a, variant)
    exp = torch.outer(torch.arange(1, cout + 1).float(), torch.arange(1, cin + 1).float())
    print("A: match", bool(torch.equal(d, exp)), "nonzero", int((d != 0).sum()), "of", d.numel())
    print("A: rows 0..2, cols 0..9:", d[0:3, 0:10].tolist())
    print("A: rows 8,9,16,32,64 col 0..3:", d[[8, 9, 16, 32, 64], 0:4].tolist())
    # B: pixel (K) pairing: gy one-hot at pixel p0 (co 0), a[ci 0][p] = flat valid index + 1
    a = torch.zeros(N, cin, H, W)
    a[0, 0] = (torch.arange(H * W).float() + 1).view(H, W)
    for (h, w) in [(0, 0), (0, 1), (0, 7), (1, 0), (1, 1), (1, 6), (2, 0), (7, 7)]:
        gy = torch.zeros(N, cout, H, W)
        gy[0, 0, h, w] = 1.0
        d = run(gy, a, variant)
        nz = [(int(i), int(j), float(d[i, j])) for i, j in (d != 0).nonzero()[:6]]
        print(f"B: gy@({h},{w}) flat {h * (W + 1) + w:3d} expect {h * W + w + 1:3d} at [0][0]: got {float(d[0, 0]):6.1f}  nonzeros {nz}")
