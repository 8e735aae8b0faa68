// tcgen05 implicit-GEMM convolution for the U-Net hot path (replaces the cuDNN conv2d calls that
// diffusers' UNet2DModel.forward makes; reference call site audiodiffusion/pipeline_audio_diffusion.py:163).
//
// GEMM view: D[pixel, cout] = sum over (segment, tap, cin) A[pixel + shift(tap), cin] * W[cout, cin, tap].
//   M = 128 output pixels (consecutive positions of the PF8 flat sequence), N = 128 output channels,
//   K = 16 input channels per tcgen05.mma.
// A operand: the PF8 layout stores 8-channel vectors of consecutive pixels contiguously, which *is* the
//   K-major no-swizzle UMMA core-matrix layout (8 rows x 16 B). A work item covers MAXG*128 consecutive flat
//   pixels; per 16 input channels ONE contiguous window per 8-channel plane (the run plus a halo of Wp+1 pixels on
//   both sides) is bulk-copied (TMA engine, UBLKCP) into shared memory, and every tap of every tile is a
//   shared-memory descriptor whose start address is shifted by (tile*128 + dh*Wp + dw) * 16 B.
//   (Measured on B200: a bulk copy costs ~130 cycles of TMA time however small it is, so few large copies.)
// Fused GroupNorm(+SiLU): the strips hold the RAW producer output; six transform warps rewrite them in place
//   (x * scale[n][c] + shift[n][c], SiLU via one tanh.approx, zero on pad/guard positions) between the TMA landing
//   and the MMA reading them, so the normalised tensor never exists in HBM (GroupNorm statistics come from the
//   producer's epilogue, scale/shift from gn_finalize_kernel).
// B operand: weights pre-packed on the device into per-(cout tile, 16-channel step, tap) 4 KB blocks; one copy.
// Residual adds are an extra 1-tap K-segment with identity weights (exact, and no epilogue loads).
// Accumulators: MAXG tiles x 128 fp32 columns in TMEM, ACC stages (see ConvCfg).
// Warp roles (12 warps): 0 bulk-copy producer, 1 MMA issuer (uniform datapath, one elected lane), 2/3/8-11 transform
//   (warp 2 also owns the TMEM allocation), 4-7 epilogue (bias + timestep embedding,
//   GroupNorm partial statistics for the consumer, bf16 store).
#include <cstdlib>

#include "conv_tc.cuh"

namespace b200ad {

constexpr int CONV_THREADS = 384;     // 12 warps
constexpr int CONV_XF_THREADS = 192;  // transform warps 2, 3, 8, 9, 10, 11

struct WorkItem {
  int n, ntile, m0, G;
};

__device__ __forceinline__ WorkItem decode_work(const ConvParams& p, int w) {
  WorkItem wi;
  wi.ntile = w % p.ntiles_n;
  const int gidx = w / p.ntiles_n;
  wi.n = gidx / p.groups_per_img;
  const int g = gidx - wi.n * p.groups_per_img;
  wi.m0 = g * (p.maxg * CONV_TM);
  const int rem = p.H * p.Wp - wi.m0;
  wi.G = min(p.maxg, (rem + CONV_TM - 1) / CONV_TM);
  return wi;
}

// one 16-byte vector (8 channels of one pixel): affine + optional SiLU in fp32, back to bf16.
// With SiLU the caller passes HALVED scale/shift: h = a/2 = x*s' + t', silu(a) = a * (0.5 + 0.5 tanh(a/2)) = h + h * tanh(h)
// -> FFMA, MUFU.TANH, FFMA per element.
template <bool SILU>
__device__ __forceinline__ uint4 xform_vec(uint4 v, const float2 (&ss)[8]) {
  uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_bf16x2(u[e]);
    float a = fmaf(f.x, ss[2 * e].x, ss[2 * e].y);
    float b = fmaf(f.y, ss[2 * e + 1].x, ss[2 * e + 1].y);
    if (SILU) {
      a = fmaf(a, tanh_approx(a), a);
      b = fmaf(b, tanh_approx(b), b);
    }
    u[e] = pack_bf16x2(a, b);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

template <int MAXG, int ACC, int STAGES>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = p.a_stage + CONV_B_STAGE;

  uint8_t* ctrl = smem + STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctrl);  // full[S], ready[S], empty[S], tmem_full[ACC], tmem_empty[ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 120);
  float* sbias = reinterpret_cast<float*>(ctrl + 128);  // 128 floats
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_ready = smem_u32(bars + STAGES);
  const uint32_t bar_empty = smem_u32(bars + 2 * STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 3 * STAGES);
  const uint32_t bar_tempty = smem_u32(bars + 3 * STAGES + ACC);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_ready + 8 * s, CONV_XF_THREADS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ producer: per stage two A windows (one per 8-channel plane) + one B block
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const WorkItem wi = decode_work(p, w);
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg& sg = p.seg[s];
        const int npix = wi.G * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr;
        const uint32_t row_bytes = (uint32_t)npix * 16u;
        const uint32_t b_bytes = (uint32_t)sg.ntaps * CONV_B_TAP;
        const uint32_t tx_bytes = 2u * row_bytes + b_bytes;
        const char* src = nullptr;   // this lane's copy: source at k-step 0, byte advance per k-step
        long long src_step = 0;
        uint32_t dst_off = 0, bytes = 0;
        if (lane < 2) {
          const int pix0 = p.lead + wi.m0 - sg.ht * p.Wp - sg.hl;
          src = reinterpret_cast<const char*>(sg.src + (long long)wi.n * sg.img_stride + ((long long)lane * p.PL + pix0) * 8);
          src_step = (long long)2 * p.PL * 16;
          dst_off = (uint32_t)lane * row_bytes;
          bytes = row_bytes;
        } else if (lane == 31) {
          src = reinterpret_cast<const char*>(sg.wpack + (long long)wi.ntile * sg.ksteps * sg.ntaps * (CONV_B_TAP / 2));
          src_step = (long long)b_bytes;
          dst_off = (uint32_t)p.a_stage;
          bytes = b_bytes;
        }
        for (int ks = 0; ks < sg.ksteps; ++ks) {
          const uint32_t full = bar_full + 8 * stage;
          if (lane == 0) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(full, tx_bytes);
          }
          __syncwarp();
          if (bytes) bulk_g2s(smem_base + stage * stage_bytes + dst_off, src, bytes, full);
          src += src_step;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer. The warp runs convergently (all operands live in uniform registers);
    // one elected lane issues. Taps outer, tiles inner: consecutive MMAs share the B descriptor and advance A / D by one tile.
    constexpr uint32_t idesc = make_idesc_bf16(CONV_TM, CONV_NT);
    constexpr uint32_t bdesc_lo_hi = ((CONV_NT / 8) * 128 >> 4) << 16;   // LBO of B
    constexpr uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, descriptor version 1
    int stage = 0;
    uint32_t phase = 0;
    int item = 0;
    const uint32_t b_off16 = (uint32_t)p.a_stage >> 4;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++item) {
      const WorkItem wi = decode_work(p, w);
      const int acc = item % ACC;
      const uint32_t d0 = tmem_base + (uint32_t)acc * (MAXG * CONV_NT);
      mbar_wait_warp(bar_tempty + 8 * acc, (((uint32_t)(item / ACC)) & 1) ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      uint32_t fresh = 1;  // first k-step of the item overwrites the accumulators
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg& sg = p.seg[s];
        const int npix = wi.G * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr;
        const uint32_t adesc_lo_hi = ((uint32_t)npix & 0x3FFF) << 16;  // LBO of A = npix * 16 B
        const int ntaps = sg.ntaps;
        for (int ks = 0; ks < sg.ksteps; ++ks) {
          mbar_wait_warp(bar_ready + 8 * stage, phase);   // strips landed and (if asked) normalised in place
          tc_fence_after();
          const uint32_t base16 = (smem_base + stage * stage_bytes) >> 4;
          if (elect_one()) {
            for (int t = 0; t < ntaps; ++t) {
              const uint32_t a_lo = (base16 + (uint32_t)sg.aoff[t]) | adesc_lo_hi;
              const uint64_t bdesc = desc_hi | (uint64_t)((base16 + b_off16 + (uint32_t)t * (CONV_B_TAP >> 4)) | bdesc_lo_hi);
              const uint32_t accum = (fresh && t == 0) ? 0u : 1u;
#pragma unroll
              for (int i = 0; i < MAXG; ++i) {
                if (i < wi.G)
                  umma_bf16(d0 + i * CONV_NT, desc_hi | (uint64_t)(a_lo + i * (CONV_TM * 16 >> 4)), bdesc, idesc, accum);
              }
            }
          }
          __syncwarp();
          fresh = 0;
          umma_commit_elect(bar_empty + 8 * stage);  // frees the stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit_elect(bar_tfull + 8 * acc);
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================ epilogue: TMEM -> regs -> (+bias,+temb) -> stats, bf16 store
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int et = threadIdx.x - 128;        // 0..127
    const long long out_img_stride = (long long)(p.cout >> 3) * p.PL * 8;
    const int hw_end = p.H * p.Wp;
    const bool do_stats = p.stats != nullptr;
    int item = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++item) {
      const WorkItem wi = decode_work(p, w);
      const int acc = item % ACC;
      const uint32_t acc_col = (uint32_t)acc * (MAXG * CONV_NT);
      {  // per-item additive vector: bias + timestep-embedding projection of this sample
        const int c = wi.ntile * CONV_NT + et;
        float b = p.bias ? p.bias[c] : 0.f;
        if (p.temb) b += p.temb[(long long)wi.n * p.temb_stride + c];
        asm volatile("bar.sync 1, 128;");   // previous item's readers are done with sbias
        sbias[et] = b;
        asm volatile("bar.sync 1, 128;");
      }
      __nv_bfloat16* out_img = p.out + (long long)wi.n * out_img_stride;
      const long long plane0 = (long long)wi.ntile * 16 * p.PL * 8;

      mbar_wait(bar_tfull + 8 * acc, ((uint32_t)(item / ACC)) & 1);
      tc_fence_after();

      float st[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 16; ++k) st[j][k] = 0.f;

      const int G = (p.dbg & 8) ? 0 : wi.G;
      for (int i = 0; i < G; ++i) {
        const int m = wi.m0 + i * CONV_TM + q * 32 + lane;
        const bool valid = (m < hw_end) && ((m % p.Wp) != p.W);
        const long long pix = (long long)(p.lead + m) * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)(i * CONV_NT + j * 32), r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]) + sbias[j * 32 + e];
          if (valid) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint4 o;
              o.x = pack_bf16x2(v[c8 * 8 + 0], v[c8 * 8 + 1]);
              o.y = pack_bf16x2(v[c8 * 8 + 2], v[c8 * 8 + 3]);
              o.z = pack_bf16x2(v[c8 * 8 + 4], v[c8 * 8 + 5]);
              o.w = pack_bf16x2(v[c8 * 8 + 6], v[c8 * 8 + 7]);
              *reinterpret_cast<uint4*>(out_img + plane0 + (long long)(j * 4 + c8) * p.PL * 8 + pix) = o;
            }
            if (do_stats) {
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
      // accumulators are drained: the MMA warp may reuse this TMEM stage
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);

      if (do_stats) {
        // warp transpose-reduce: 16 values per 32-column chunk -> one lane per value, fp64 atomics
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
  } else {
    // ================================ transform warps (2, 3, 8..11): GroupNorm(+SiLU) of the landed strips, in place
    const int tt = ((warp < 4) ? (warp - 2) : (warp - 6)) * 32 + lane;  // 0..191
    const int hw_end = p.H * p.Wp;
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const WorkItem wi = decode_work(p, w);
      for (int s = 0; s < p.nseg; ++s) {
        const ConvSeg& sg = p.seg[s];
        const int npix = wi.G * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr;
        const float2* ssn = sg.ss ? sg.ss + (long long)wi.n * sg.ss_stride : nullptr;
        // flat position of this thread's first pixel; (row, col) advance incrementally (128 pixels per step)
        const int m_first = wi.m0 - sg.ht * p.Wp - sg.hl + tt;
        int row0 = (m_first >= 0) ? m_first / p.Wp : -1 - ((-1 - m_first) / p.Wp);  // floor division
        int col0 = m_first - row0 * p.Wp;
        const int drow = CONV_XF_THREADS / p.Wp, dcol = CONV_XF_THREADS - drow * p.Wp;
        const bool silu = sg.silu != 0;
        for (int ks = 0; ks < sg.ksteps; ++ks) {
          float2 ss0[8], ss1[8];
          if (ssn) {
            const float4* sp = reinterpret_cast<const float4*>(ssn + ks * 16);
            const float hs = silu ? 0.5f : 1.0f;  // SiLU path works on a/2 (see xform_vec)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 a = __ldg(sp + e), b = __ldg(sp + 4 + e);
              ss0[2 * e] = make_float2(a.x * hs, a.y * hs); ss0[2 * e + 1] = make_float2(a.z * hs, a.w * hs);
              ss1[2 * e] = make_float2(b.x * hs, b.y * hs); ss1[2 * e + 1] = make_float2(b.z * hs, b.w * hs);
            }
          }
          mbar_wait(bar_full + 8 * stage, phase);
          if (ssn && !(p.dbg & 64)) {
            uint4* base = reinterpret_cast<uint4*>(smem + stage * stage_bytes);
            int row = row0, col = col0;
            for (int px = tt; px < npix; px += CONV_XF_THREADS) {
              const bool valid = (row >= 0) && (row < p.H) && (col < p.W);
              uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
              if (valid) {
                a = base[px];
                b = base[npix + px];
                if (silu && !(p.dbg & 256)) { a = xform_vec<true>(a, ss0); b = xform_vec<true>(b, ss1); }
                else      { a = xform_vec<false>(a, ss0); b = xform_vec<false>(b, ss1); }
              }
              base[px] = a;
              base[npix + px] = b;
              row += drow; col += dcol;
              if (col >= p.Wp) { col -= p.Wp; ++row; }
            }
            if (!(p.dbg & 128)) fence_proxy_async_smem();
          }
          mbar_arrive(bar_ready + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    (void)hw_end;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int CFG>
static cudaError_t launch_cfg(const ConvParams& p, int grid, size_t smem, cudaStream_t stream) {
  constexpr ConvCfg c = CONV_CFGS[CFG];
  auto kern = conv_tc_kernel<c.maxg, c.acc, c.stages>;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = smem;
  }
  kern<<<grid, CONV_THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv_tc(const ConvParams& p_in, int num_sms, cudaStream_t stream) {
  static int dbg = -1, cfg_env = -1;
  if (dbg < 0) {
    const char* e = getenv("B200AD_CONV_DBG");
    dbg = e ? atoi(e) : 0;
    const char* c = getenv("B200AD_CONV_CFG");
    cfg_env = c ? atoi(c) : 0;  // measured: with the GroupNorm transform fused, cfg 0 (less halo per MMA) wins
    if (cfg_env < 0 || cfg_env > 1) cfg_env = 0;
  }
  ConvParams p = p_in;
  p.dbg = dbg;
  const ConvCfg cfg = CONV_CFGS[cfg_env];
  p.maxg = cfg.maxg;
  p.groups_per_img = (p.H * p.Wp + cfg.maxg * CONV_TM - 1) / (cfg.maxg * CONV_TM);
  p.ntiles_n = p.cout / CONV_NT;
  p.total_work = p.N * p.groups_per_img * p.ntiles_n;
  // the two A windows of one k-step must fit the stage buffer
  int a_stage = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const ConvSeg& sg = p.seg[s];
    const int npix = cfg.maxg * CONV_TM + (sg.ht + sg.hb) * p.Wp + sg.hl + sg.hr;
    a_stage = npix * 32 > a_stage ? npix * 32 : a_stage;
    if (sg.ntaps > CONV_MAXTAPS || npix > 0x3FFF) return cudaErrorInvalidValue;
    for (int t = 0; t < sg.ntaps; ++t) p.seg[s].aoff[t] = (sg.dh[t] + sg.ht) * p.Wp + sg.dw[t] + sg.hl;
  }
  p.a_stage = (a_stage + 255) & ~255;
  const size_t smem = (size_t)cfg.stages * (p.a_stage + CONV_B_STAGE) + 1024;
  if (smem > (size_t)CONV_SMEM_MAX) return cudaErrorInvalidValue;  // image too wide for this tiling
  const int grid = p.total_work < num_sms ? p.total_work : num_sms;
  if (grid <= 0) return cudaSuccess;
  return cfg_env == 0 ? launch_cfg<0>(p, grid, smem, stream) : launch_cfg<1>(p, grid, smem, stream);
}

// ------------------------------------------------------------------------------------ identity weights
__global__ void pack_identity_kernel(int channels, __nv_bfloat16* __restrict__ dst, long long nvec) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nvec) return;
  const int r = (int)(id & 7), n8 = (int)((id >> 3) & 15), k8 = (int)((id >> 7) & 1);
  const long long rest = id >> 8;
  const int ksteps = channels / 16;
  const int ks = (int)(rest % ksteps), ntile = (int)(rest / ksteps);
  const int co = ntile * 128 + n8 * 8 + r;
  const int ci0 = ks * 16 + k8 * 8;
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(co == ci0 + 2 * e ? 1.f : 0.f, co == ci0 + 2 * e + 1 ? 1.f : 0.f);
  reinterpret_cast<uint4*>(dst)[id] = make_uint4(o[0], o[1], o[2], o[3]);
}
cudaError_t launch_pack_identity(int channels, __nv_bfloat16* dst, cudaStream_t s) {
  const long long nvec = (long long)(channels / 128) * (channels / 16) * 256;
  pack_identity_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, s>>>(channels, dst, nvec);
  return cudaGetLastError();
}

}  // namespace b200ad
