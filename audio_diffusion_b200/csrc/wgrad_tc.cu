// Weight gradient of a convolution on tcgen05 (DESIGN.md §3 "Training step").
//   dW[co][ci][tap] = sum over samples and flat pixels p of  gY[co][p] * A[ci][p + shift(tap)]
// GEMM view per tap: D[M = 128 co][N = NCI ci] += sum_K gY[co][K = pixel] * A[ci][K = pixel + shift]: the reduction runs over
// PIXELS, so both operands are "MN-major": in PF8 the 8 channels of a pixel are contiguous (MN) and consecutive pixels are
// 16 B apart (K) — a run of 8 pixels of one 8-channel plane IS the 8x8 MN-major core matrix (128 B): LBO (next 8 pixels) =
// 128 B, SBO (next 8 channels) = the plane pitch of the staged window (verified on hardware: tools/wgrad_probe.py).
// Work decomposition: one CTA = (128-co tile, NCI-ci tile, ONE ROW of taps (same dh, dw in {-1,0,+1}), a share of the pixel
// blocks).  A tap row needs only a 1-pixel halo (the row offset dh*Wp moves the window start), its <= 3 taps are shifted
// descriptors of one activation window, and N = 128 keeps the per-MMA fixed cost small (measured: N = 32 MMAs cost ~110
// cycles each, 7x their tensor time).  The gY block (A operand) is latched in the A collector across the taps of a row.
// TMEM: 3 taps x NCI fp32 columns.  Partial dW is added to global memory with fp32 atomics (split-K over pixels and CTAs).
#include <cstdlib>

#include "bwd_kernels.cuh"
#include "conv_tc.cuh"

namespace b200ad {

struct WgradParams {
  const __nv_bfloat16* gy;    // PF8 view, cout channels
  const __nv_bfloat16* act;   // PF8 view, cin channels (the conv's input as the forward pass saw it)
  float* dw;                  // fp32, element (co, ci, t) at (co * cin_total + ci_off + ci) * ntaps_total + tapidx
  int N, H, W, Wp, lead, PL, cin, cout;
  int gy_img_planes, act_img_planes, cin_total, ci_off, ntaps_total;
  int nrows;                  // tap rows (distinct dh)
  int row_off[3];             // dh * Wp of the row
  int row_ntaps[3];
  int row_dw[3][3];           // dw of each tap of the row
  int row_tapidx[3][3];       // destination tap index
  int P;                      // pixels per staged block (multiple of 16)
  int nblk;                   // blocks per image
};

constexpr int WG_THREADS = 192;   // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

template <int NCI, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gy_bytes = p.P * 256;                             // 16 planes x P pixels x 16 B
  const int act_pix = p.P + 2;                                // 1-pixel halo on both sides
  const int act_bytes = (NCI / 8) * act_pix * 16;
  const int stage_bytes = (gy_bytes + act_bytes + 127) & ~127;
  uint8_t* ctrl = smem + STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);         // full[S], empty[S], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 128);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + STAGES), bar_done = smem_u32(bars + 2 * STAGES);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int row = blockIdx.y % p.nrows, cj = blockIdx.y / p.nrows, ct = blockIdx.z;
  const int total_blocks = p.N * p.nblk;
  const int ntaps = p.row_ntaps[row];

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int roff = p.row_off[row];
      for (int b = blockIdx.x; b < total_blocks; b += gridDim.x) {
        const int n = b / p.nblk, m0 = (b - n * p.nblk) * p.P;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)(gy_bytes + act_bytes));
        const uint32_t dst = smem_base + stage * stage_bytes;
        for (int pl = 0; pl < 16; ++pl) {
          const __nv_bfloat16* src = p.gy + (((long long)n * p.gy_img_planes + ct * 16 + pl) * p.PL + p.lead + m0) * 8;
          bulk_g2s(dst + pl * p.P * 16, src, (uint32_t)p.P * 16u, bar_full + 8 * stage);
        }
        for (int pl = 0; pl < NCI / 8; ++pl) {
          const __nv_bfloat16* src =
              p.act + (((long long)n * p.act_img_planes + cj * (NCI / 8) + pl) * p.PL + p.lead + m0 + roff - 1) * 8;
          bulk_g2s(dst + gy_bytes + pl * act_pix * 16, src, (uint32_t)act_pix * 16u, bar_full + 8 * stage);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the warp runs convergently (operands in uniform registers, waits exit on a vote), one elected lane issues.
    constexpr uint32_t idesc = make_idesc_bf16(128, NCI) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint64_t a_hi = (uint64_t)(((uint32_t)p.P & 0x3FFF) | (1u << 14)) << 32;    // SBO = P * 16 B, descriptor version 1
    const uint64_t b_hi = (uint64_t)(((uint32_t)act_pix & 0x3FFF) | (1u << 14)) << 32;
    constexpr uint32_t lbo = (128u >> 4) << 16;
    const int dw0 = p.row_dw[row][0] + 1, dw1 = p.row_dw[row][1] + 1, dw2 = p.row_dw[row][2] + 1;   // + halo
    int stage = 0;
    uint32_t phase = 0;
    uint32_t first = 1;
    for (int b = blockIdx.x; b < total_blocks; b += gridDim.x) {
      mbar_wait_warp(bar_full + 8 * stage, phase);
      tc_fence_after();
      const uint32_t a16 = (smem_base + stage * stage_bytes) >> 4;
      const uint32_t b16 = a16 + ((uint32_t)gy_bytes >> 4);
      const int ksteps = p.P / 16;
      if (elect_one()) {
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adesc = a_hi | (uint64_t)((a16 + (uint32_t)k * 16u) | lbo);
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          const uint32_t bk = b16 + (uint32_t)k * 16u;
          const uint64_t bd0 = b_hi | (uint64_t)((bk + (uint32_t)dw0) | lbo);
          if (ntaps == 1) {
            umma_bf16(tmem_base, adesc, bd0, idesc, acc);
          } else {
            umma_bf16_afill(tmem_base, adesc, bd0, idesc, acc);
            const uint64_t bd1 = b_hi | (uint64_t)((bk + (uint32_t)dw1) | lbo);
            if (ntaps == 2) {
              umma_bf16_alast(tmem_base + NCI, adesc, bd1, idesc, acc);
            } else {
              umma_bf16_ause(tmem_base + NCI, adesc, bd1, idesc, acc);
              const uint64_t bd2 = b_hi | (uint64_t)((bk + (uint32_t)dw2) | lbo);
              umma_bf16_alast(tmem_base + 2 * NCI, adesc, bd2, idesc, acc);
            }
          }
        }
      }
      __syncwarp();
      first = 0;
      umma_commit_elect(bar_empty + 8 * stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_elect(bar_done);
  } else {
    // epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 = output channels; columns = the CTA's input channels per tap
    const int q = warp & 3;
    if (blockIdx.x < total_blocks) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      const int co = ct * 128 + q * 32 + lane;
      for (int t = 0; t < ntaps; ++t) {
        const int tapidx = p.row_tapidx[row][t];
        for (int c0 = 0; c0 < NCI; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NCI + c0), r);
          tmem_ld_wait();
          float* dst = p.dw + ((long long)co * p.cin_total + p.ci_off + cj * NCI + c0) * p.ntaps_total + tapidx;
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(dst + (long long)e * p.ntaps_total, __uint_as_float(r[e]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int NCI, int STAGES>
static cudaError_t launch_wg(const WgradParams& p, int num_sms, cudaStream_t s) {
  const int act_pix = p.P + 2;
  const int stage_bytes = (p.P * 256 + (NCI / 8) * act_pix * 16 + 127) & ~127;
  const size_t smem = (size_t)STAGES * stage_bytes + 256;
  if (smem > (size_t)CONV_SMEM_MAX) return cudaErrorInvalidValue;
  auto kern = wgrad_tc_kernel<NCI, STAGES>;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  const int tiles = p.nrows * (p.cin / NCI) * (p.cout / 128);
  int splits = num_sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.N * p.nblk) splits = p.N * p.nblk;
  dim3 grid(splits, p.nrows * (p.cin / NCI), p.cout / 128);
  kern<<<grid, WG_THREADS, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_wgrad_tc(const WgradDesc& d, int num_sms, cudaStream_t s) {
  if (d.cout % 128 || d.cin % 32 || d.ntaps < 1 || d.ntaps > 9) return cudaErrorInvalidValue;
  const Geom g = make_geom(d.N, d.H, d.W);
  WgradParams p{};
  p.gy = d.gy; p.act = d.act; p.dw = d.dw;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Wp = g.Wp; p.lead = g.lead; p.PL = g.PL; p.cin = d.cin; p.cout = d.cout;
  p.gy_img_planes = d.gy_img_planes; p.act_img_planes = d.act_img_planes;
  p.cin_total = d.cin_total; p.ci_off = d.ci_off; p.ntaps_total = d.ntaps_total;
  // group the taps into rows of equal dh
  p.nrows = 0;
  for (int t = 0; t < d.ntaps; ++t) {
    if (d.dh[t] < -1 || d.dh[t] > 1 || d.dw_[t] < -1 || d.dw_[t] > 1) return cudaErrorInvalidValue;
    int r = -1;
    for (int k = 0; k < p.nrows; ++k)
      if (p.row_off[k] == d.dh[t] * g.Wp) r = k;
    if (r < 0) {
      if (p.nrows == 3) return cudaErrorInvalidValue;
      r = p.nrows++;
      p.row_off[r] = d.dh[t] * g.Wp;
      p.row_ntaps[r] = 0;
    }
    if (p.row_ntaps[r] == 3) return cudaErrorInvalidValue;
    p.row_dw[r][p.row_ntaps[r]] = d.dw_[t];
    p.row_tapidx[r][p.row_ntaps[r]] = d.tapidx[t];
    ++p.row_ntaps[r];
  }
  const int flat = d.H * g.Wp;
  if (d.cin % 128 == 0) {
    // 128-wide ci tiles: 192-pixel stages, two of them (2 x 97 KB)
    p.P = flat <= 128 ? 128 : 192;
    p.nblk = (flat + p.P - 1) / p.P;
    return launch_wg<128, 2>(p, num_sms, s);
  }
  p.P = flat <= 128 ? 128 : 256;
  p.nblk = (flat + p.P - 1) / p.P;
  return launch_wg<32, 2>(p, num_sms, s);
}

}  // namespace b200ad
