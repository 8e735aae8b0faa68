// Weight gradient of a stride-1 conv on tcgen05 (first, correctness-first version; DESIGN.md §6 has the roofline analysis).
//   dW[co][ci][tap] = sum over samples and flat pixels p of  gY[co][p] * A[ci][p + shift(tap)]
// GEMM view per tap: D[M = 128 co][N = 32 ci] += sum_K gY[co][K = pixel] * A[ci][K = pixel + shift]: the reduction runs over
// PIXELS, so both operands are "MN-major": in PF8 the 8 channels of a pixel are contiguous (MN) and consecutive pixels are
// 16 B apart (K) — a run of 8 pixels of one 8-channel plane IS the 8x8 MN-major core matrix (128 B), planes are SBO apart.
// The tap shift is a shifted start address of the activation window, exactly as in the forward kernel.
// TMEM: 9 taps x 32 fp32 columns per CTA; pixel range split across CTAs, partial dW added to global memory with fp32 atomics.
#include <cstdlib>

#include "bwd_kernels.cuh"
#include "conv_tc.cuh"

namespace b200ad {

struct WgradParams {
  const __nv_bfloat16* gy;    // PF8 view, cout channels
  const __nv_bfloat16* act;   // PF8 view, cin channels (the conv's input as the forward pass saw it)
  float* dw;                  // fp32, element (co, ci, t) at (co * cin_total + ci_off + ci) * ntaps_total + tapidx[t]
  int N, H, W, Wp, lead, PL, cin, cout, ntaps;
  int gy_img_planes, act_img_planes, cin_total, ci_off, ntaps_total;
  int shift[9];               // dh * Wp + dw per tap
  int tapidx[9];
  int halo;                   // max |shift|
  int P;                      // pixels per staged block (multiple of 16)
  int nblk;                   // blocks per image
};

constexpr int WG_NCI = 32;        // ci per CTA (N of the MMA)
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 192;   // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gy_bytes = p.P * 256;                             // 16 planes x P pixels x 16 B
  const int act_pix = p.P + 2 * p.halo;
  const int act_bytes = (WG_NCI / 8) * act_pix * 16;
  const int stage_bytes = (gy_bytes + act_bytes + 127) & ~127;
  uint8_t* ctrl = smem + WG_STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);         // full[S], empty[S], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 64);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + WG_STAGES), bar_done = smem_u32(bars + 2 * WG_STAGES);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int cj = blockIdx.y, ct = blockIdx.z;
  const int total_blocks = p.N * p.nblk;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = blockIdx.x; b < total_blocks; b += gridDim.x) {
        const int n = b / p.nblk, m0 = (b - n * p.nblk) * p.P;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)(gy_bytes + act_bytes));
        const uint32_t dst = smem_base + stage * stage_bytes;
        for (int pl = 0; pl < 16; ++pl) {
          const __nv_bfloat16* src = p.gy + (((long long)n * p.gy_img_planes + ct * 16 + pl) * p.PL + p.lead + m0) * 8;
          bulk_g2s(dst + pl * p.P * 16, src, (uint32_t)p.P * 16u, bar_full + 8 * stage);
        }
        for (int pl = 0; pl < WG_NCI / 8; ++pl) {
          const __nv_bfloat16* src =
              p.act + (((long long)n * p.act_img_planes + cj * (WG_NCI / 8) + pl) * p.PL + p.lead + m0 - p.halo) * 8;
          bulk_g2s(dst + gy_bytes + pl * act_pix * 16, src, (uint32_t)act_pix * 16u, bar_full + 8 * stage);
        }
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the warp runs convergently (operands in uniform registers, waits exit on a vote), one elected lane issues.
    // Per 16 pixels: the gY block (A, 128 co x 16 px) is latched in the A collector and reused by all taps.
    const uint32_t idesc = make_idesc_bf16(128, WG_NCI) | (1u << 15) | (1u << 16);   // A and B MN-major
    // MN-major, no swizzle: LBO = distance of the next 8 pixels (K) = 128 B, SBO = distance of the next 8 channels (MN) =
    // one plane of the staged window  [verified on hardware: tools/wgrad_probe.py]
    const uint64_t a_hi = (uint64_t)(((uint32_t)p.P & 0x3FFF) | (1u << 14)) << 32;          // SBO = P * 16 B, version 1
    const uint64_t b_hi = (uint64_t)(((uint32_t)act_pix & 0x3FFF) | (1u << 14)) << 32;
    constexpr uint32_t lbo = (128u >> 4) << 16;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t first = 1;
    for (int b = blockIdx.x; b < total_blocks; b += gridDim.x) {
      mbar_wait_warp(bar_full + 8 * stage, phase);
      tc_fence_after();
      const uint32_t a16 = (smem_base + stage * stage_bytes) >> 4;
      const uint32_t b16 = a16 + ((uint32_t)gy_bytes >> 4) + (uint32_t)p.halo;
      const int ksteps = p.P / 16, ntaps = p.ntaps;
      if (elect_one()) {
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = a_hi | (uint64_t)((a16 + (uint32_t)k * 16u) | lbo);
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          for (int t = 0; t < ntaps; ++t) {
            const uint64_t bdesc = b_hi | (uint64_t)((b16 + (uint32_t)(k * 16 + p.shift[t])) | lbo);
            const uint32_t d = tmem_base + (uint32_t)t * WG_NCI;
            if (ntaps == 1) umma_bf16(d, adesc, bdesc, idesc, acc);
            else if (t == 0) umma_bf16_afill(d, adesc, bdesc, idesc, acc);
            else if (t == ntaps - 1) umma_bf16_alast(d, adesc, bdesc, idesc, acc);
            else umma_bf16_ause(d, adesc, bdesc, idesc, acc);
          }
        }
      }
      __syncwarp();
      first = 0;
      umma_commit_elect(bar_empty + 8 * stage);
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_elect(bar_done);
  } else {
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 = output channels; 32 columns = the CTA's input channels
    const int q = warp & 3;
    const bool has_work = blockIdx.x < total_blocks;
    if (has_work) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int co = ct * 128 + q * 32 + lane;
      for (int t = 0; t < p.ntaps; ++t) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * WG_NCI, r);
        tmem_ld_wait();
        float* dst = p.dw + ((long long)co * p.cin_total + p.ci_off + cj * WG_NCI) * p.ntaps_total + p.tapidx[t];
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(dst + (long long)e * p.ntaps_total, __uint_as_float(r[e]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

cudaError_t launch_wgrad_tc(const WgradDesc& d, int num_sms, cudaStream_t s) {
  if (d.cout % 128 || d.cin % WG_NCI || d.ntaps < 1 || d.ntaps > 9) return cudaErrorInvalidValue;
  const Geom g = make_geom(d.N, d.H, d.W);
  WgradParams p{};
  p.gy = d.gy; p.act = d.act; p.dw = d.dw;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Wp = g.Wp; p.lead = g.lead; p.PL = g.PL; p.cin = d.cin; p.cout = d.cout;
  p.ntaps = d.ntaps;
  p.gy_img_planes = d.gy_img_planes; p.act_img_planes = d.act_img_planes;
  p.cin_total = d.cin_total; p.ci_off = d.ci_off; p.ntaps_total = d.ntaps_total;
  p.halo = 0;
  for (int t = 0; t < d.ntaps; ++t) {
    p.shift[t] = d.dh[t] * g.Wp + d.dw_[t];
    p.tapidx[t] = d.tapidx[t];
    const int a = p.shift[t] < 0 ? -p.shift[t] : p.shift[t];
    if (a > p.halo) p.halo = a;
  }
  if (p.halo > g.lead - 1) return cudaErrorInvalidValue;   // the window may not start before the plane
  // pixels per staged block: 256 when two stages fit (fewer, larger bulk copies per pixel), else 128
  p.P = 256;
  if ((size_t)WG_STAGES * ((256 * 256 + (WG_NCI / 8) * (256 + 2 * p.halo) * 16 + 127) & ~127) + 256 > (size_t)CONV_SMEM_MAX ||
      d.H * g.Wp <= 128)
    p.P = 128;
  p.nblk = (d.H * g.Wp + p.P - 1) / p.P;
  const int act_pix = p.P + 2 * p.halo;
  const int stage_bytes = (p.P * 256 + (WG_NCI / 8) * act_pix * 16 + 127) & ~127;
  const size_t smem = (size_t)WG_STAGES * stage_bytes + 256;
  if (smem > (size_t)CONV_SMEM_MAX) return cudaErrorInvalidValue;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  const int tiles = (d.cin / WG_NCI) * (d.cout / 128);
  int splits = num_sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > d.N * p.nblk) splits = d.N * p.nblk;
  dim3 grid(splits, d.cin / WG_NCI, d.cout / 128);
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace b200ad
