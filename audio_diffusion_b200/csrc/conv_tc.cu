// tcgen05 implicit-GEMM convolution for the U-Net hot path (replaces the cuDNN conv2d calls that
// diffusers' UNet2DModel.forward makes; reference call site audiodiffusion/pipeline_audio_diffusion.py:163).
//
// GEMM view: D[pixel, cout] = sum over (segment, tap, cin) A[pixel + shift(tap), cin] * W[cout, cin, tap].
//   M = 128 output pixels (consecutive positions of the PF8 flat sequence), N = 128 output channels,
//   K = 16 input channels per tcgen05.mma.
// A operand: the PF8 layout stores 8-channel vectors of consecutive pixels contiguously, which *is* the
//   K-major no-swizzle UMMA core-matrix layout (8 rows x 16 B). One strip of (128 + halo) pixels per image
//   row is bulk-copied (TMA engine) into shared memory once per 16 channels and every 3x3 tap is issued as a
//   shared-memory descriptor whose start address is shifted by (dw + halo) * 16 B: 9 taps, one load.
// B operand: weights pre-packed on the device into per-(cout tile, 16-channel step, tap) 4 KB blocks.
// Accumulators: up to 4 pixel tiles x 128 fp32 columns = the whole TMEM, so one weight stage feeds 4 tiles.
// Warp roles: warp 0 bulk-copy producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue
//   (bias + timestep-embedding + residual, GroupNorm partial statistics for the consumer, bf16 store).
#include <cstdlib>

#include "conv_tc.cuh"

namespace b200ad {

struct WorkItem {
  int n, ntile, m0, G, tile_stride;
};

__device__ __forceinline__ WorkItem decode_work(const ConvParams& p, int w) {
  WorkItem wi;
  wi.ntile = w % p.ntiles_n;
  int gidx = w / p.ntiles_n;
  wi.n = gidx / p.groups_per_img;
  int g = gidx - wi.n * p.groups_per_img;
  if (p.wide) {
    int gw = p.W >> 7;
    int rg = g / gw, cb = g - rg * gw;
    wi.m0 = rg * CONV_MAXG * p.Wp + cb * CONV_TM;
    wi.G = min(CONV_MAXG, p.H - rg * CONV_MAXG);
    wi.tile_stride = p.Wp;
  } else {
    wi.m0 = g * (CONV_MAXG * CONV_TM);
    int rem = p.H * p.Wp - wi.m0;
    wi.G = min(CONV_MAXG, (rem + CONV_TM - 1) / CONV_TM);
    wi.tile_stride = CONV_TM;
  }
  return wi;
}

__global__ void __launch_bounds__(256, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint8_t* ctrl = smem + CONV_STAGES * CONV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);           // full[3], empty[3], tmem_full, tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 64);
  float* sbias = reinterpret_cast<float*>(ctrl + 128);          // 128 floats
  uint2* mtab = reinterpret_cast<uint2*>(ctrl + 640);           // 36 (tile, tap) descriptor-offset entries
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + CONV_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * CONV_STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 2 * CONV_STAGES + 1);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CONV_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 128);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ producer: bulk copies of A strips and B weight blocks.
    // Lane 0 owns the barrier protocol; lanes 0..2*nwin-1 each issue one A copy, lane 31 the B copy.
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const WorkItem wi = decode_work(p, w);
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg& sg = p.seg[s];
        const int nwin = p.wide ? (wi.G + sg.ht + sg.hb) : 1;
        const int npix = p.wide ? (CONV_TM + sg.hl + sg.hr)
                                : (wi.G * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr);
        const uint32_t row_bytes = (uint32_t)npix * 16u;
        const uint32_t b_bytes = (uint32_t)sg.ntaps * CONV_B_TAP;
        const uint32_t tx_bytes = (uint32_t)nwin * 2u * row_bytes + b_bytes;
        // this lane's copy: source at k-step 0 and byte advance per k-step
        const char* src = nullptr;
        long long src_step = 0;
        uint32_t dst_off = 0, bytes = 0;
        if (lane < 2 * nwin) {
          const int r = lane >> 1, pl = lane & 1;
          const int pix0 = p.lead + wi.m0 + (r - sg.ht) * p.Wp - sg.hl;
          src = reinterpret_cast<const char*>(sg.src + (long long)wi.n * sg.img_stride + ((long long)pl * p.PL + pix0) * 8);
          src_step = (long long)2 * p.PL * 16;
          dst_off = (uint32_t)lane * row_bytes;
          bytes = row_bytes;
        } else if (lane == 31) {
          src = reinterpret_cast<const char*>(sg.wpack + (long long)wi.ntile * sg.ksteps * sg.ntaps * (CONV_B_TAP / 2));
          src_step = (long long)b_bytes;
          dst_off = CONV_A_STAGE;
          bytes = b_bytes;
        }
        for (int ks = 0; ks < sg.ksteps; ++ks) {
          const uint32_t full = bar_full + 8 * stage;
          if (lane == 0) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(full, tx_bytes);
          }
          __syncwarp();
          if (bytes) bulk_g2s(smem_base + stage * CONV_STAGE_BYTES + dst_off, src, bytes, full);
          src += src_step;
          if (++stage == CONV_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer. The whole warp builds a per-segment table of descriptor offsets
    // (one entry per (tile, tap)); lane 0 then issues one tcgen05.mma per entry with two shared loads and a few ALU ops.
    constexpr uint32_t idesc = make_idesc_bf16(CONV_TM, CONV_NT);
    constexpr uint32_t bdesc_lo_hi = ((CONV_NT / 8) * 128 >> 4) << 16;   // LBO of B
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);               // SBO = 128 B, descriptor version 1
    int stage = 0;
    uint32_t phase = 0, tphase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const WorkItem wi = decode_work(p, w);
      bool first = true;
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg& sg = p.seg[s];
        const int npix = p.wide ? (CONV_TM + sg.hl + sg.hr)
                                : (wi.G * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr);
        const int cnt = wi.G * sg.ntaps;
        __syncwarp();
        for (int k = lane; k < cnt; k += 32) {
          const int i = k / sg.ntaps, t = k - i * sg.ntaps;
          uint32_t a_off;
          if (p.wide)
            a_off = (uint32_t)((i + sg.dh[t] + sg.ht) * 2 * npix + sg.dw[t] + sg.hl);          // 16-byte units
          else
            a_off = (uint32_t)(i * CONV_TM + (sg.dh[t] + sg.ht) * p.Wp + sg.dw[t] + sg.hl);
          mtab[k] = make_uint2(a_off, ((uint32_t)(CONV_A_STAGE + t * CONV_B_TAP) >> 4) | ((uint32_t)i << 16) |
                                          (t == 0 ? 0x80000000u : 0u));
        }
        __syncwarp();
        if (lane == 0) {
          const uint32_t adesc_lo_hi = ((uint32_t)npix & 0x3FFF) << 16;  // LBO of A = npix * 16 B
          if (s == 0) {
            mbar_wait(bar_tempty, tphase ^ 1);  // epilogue has drained the accumulators of the previous item
            tc_fence_after();
          }
          for (int ks = 0; ks < sg.ksteps; ++ks) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t base16 = (smem_base + stage * CONV_STAGE_BYTES) >> 4;
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
              const uint2 e = mtab[k];
              const uint64_t adesc = ((uint64_t)desc_hi << 32) | (((base16 + e.x) & 0x3FFF) | adesc_lo_hi);
              const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (((base16 + (e.y & 0xFFFF)) & 0x3FFF) | bdesc_lo_hi);
              const uint32_t d = tmem_base + ((e.y >> 16) & 0x7) * CONV_NT;
              umma_bf16(d, adesc, bdesc, idesc, (first && (e.y & 0x80000000u)) ? 0u : 1u);
            }
            first = false;
            umma_commit(bar_empty + 8 * stage);  // frees the stage when these MMAs retire
            if (++stage == CONV_STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          for (int ks = 0; ks < sg.ksteps; ++ks)
            if (++stage == CONV_STAGES) { stage = 0; phase ^= 1; }
          first = false;
        }
      }
      if (lane == 0) umma_commit(bar_tfull);
      tphase ^= 1;
    }
  } else if (warp >= 4) {
    // ================================ epilogue: TMEM -> regs -> (+bias,+temb,+residual) -> stats, bf16 store
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int et = threadIdx.x - 128;        // 0..127
    uint32_t tphase = 0;
    const long long out_img_stride = (long long)(p.cout >> 3) * p.PL * 8;
    const int hw_end = p.H * p.Wp;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const WorkItem wi = decode_work(p, w);
      // per-item additive vector: bias + timestep-embedding projection of this sample
      {
        const int c = wi.ntile * CONV_NT + et;
        float b = p.bias ? p.bias[c] : 0.f;
        if (p.temb) b += p.temb[(long long)wi.n * p.temb_stride + c];
        asm volatile("bar.sync 1, 128;");   // previous item's readers are done with sbias
        sbias[et] = b;
        asm volatile("bar.sync 1, 128;");
      }
      mbar_wait(bar_tfull, tphase);
      tc_fence_after();
      tphase ^= 1;

      float st[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 16; ++k) st[j][k] = 0.f;

      __nv_bfloat16* out_img = p.out + (long long)wi.n * out_img_stride;
      const __nv_bfloat16* res_img = p.res ? p.res + (long long)wi.n * out_img_stride : nullptr;
      for (int i = 0; i < ((p.dbg & 8) ? 0 : wi.G); ++i) {
        const int m = wi.m0 + i * wi.tile_stride + q * 32 + lane;
        bool valid = true;
        if (!p.wide) valid = (m < hw_end) && ((m % p.Wp) != p.W);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t r[32];
          if (!(p.dbg & 4)) {
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * CONV_NT + j * 32), r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = 0;
          }
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) + sbias[j * 32 + e];
          if (valid) {
            const long long pix = (long long)(p.lead + m) * 8;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              const long long off = (long long)((wi.ntile * 16 + j * 4 + c8)) * p.PL * 8 + pix;
              if (res_img) {
                const uint4 rv = *reinterpret_cast<const uint4*>(res_img + off);
                float2 f;
                f = unpack_bf16x2(rv.x); v[c8 * 8 + 0] += f.x; v[c8 * 8 + 1] += f.y;
                f = unpack_bf16x2(rv.y); v[c8 * 8 + 2] += f.x; v[c8 * 8 + 3] += f.y;
                f = unpack_bf16x2(rv.z); v[c8 * 8 + 4] += f.x; v[c8 * 8 + 5] += f.y;
                f = unpack_bf16x2(rv.w); v[c8 * 8 + 6] += f.x; v[c8 * 8 + 7] += f.y;
              }
              uint4 o;
              o.x = pack_bf16x2(v[c8 * 8 + 0], v[c8 * 8 + 1]);
              o.y = pack_bf16x2(v[c8 * 8 + 2], v[c8 * 8 + 3]);
              o.z = pack_bf16x2(v[c8 * 8 + 4], v[c8 * 8 + 5]);
              o.w = pack_bf16x2(v[c8 * 8 + 6], v[c8 * 8 + 7]);
              if (!(p.dbg & 2)) *reinterpret_cast<uint4*>(out_img + off) = o;
            }
            if (p.stats && !(p.dbg & 1)) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float a = v[4 * k], b = v[4 * k + 1], c = v[4 * k + 2], d = v[4 * k + 3];
                st[j][k] += (a + b) + (c + d);
                st[j][8 + k] += (a * a + b * b) + (c * c + d * d);
              }
            }
          }
        }
      }
      // accumulators are drained: let the MMA warp start the next item
      tc_fence_before();
      mbar_arrive(bar_tempty);

      if (p.stats && !(p.dbg & 1)) {
        // warp transpose-reduce: 16 values per 32-column chunk -> one lane per value
        stat_t* sdst = p.stats + ((long long)wi.n * (p.cout >> 2) + wi.ntile * 32) * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v16[k] = st[j][k];
#pragma unroll
          for (int half = 8, sh = 16; half >= 1; half >>= 1, sh >>= 1) {
            const bool upper = (lane & sh) != 0;
#pragma unroll
            for (int k = 0; k < half; ++k) {
              const float lo = v16[k], hi = v16[k + half];
              const float send = upper ? lo : hi;
              const float keep = upper ? hi : lo;
              v16[k] = keep + __shfl_xor_sync(0xffffffffu, send, sh);
            }
          }
          float tot = v16[0] + __shfl_xor_sync(0xffffffffu, v16[0], 1);
          if ((lane & 1) == 0) {
            const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            const int is_sq = idx >> 3, quad = idx & 7;
            atomicAdd(sdst + (j * 8 + quad) * 2 + is_sq, (stat_t)tot);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

cudaError_t launch_conv_tc(const ConvParams& p_in, int num_sms, cudaStream_t stream) {
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("B200AD_CONV_DBG");
    dbg = e ? atoi(e) : 0;
  }
  ConvParams p = p_in;
  p.dbg = dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int grid = p.total_work < num_sms ? p.total_work : num_sms;
  if (grid <= 0) return cudaSuccess;
  conv_tc_kernel<<<grid, 256, CONV_SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200ad
