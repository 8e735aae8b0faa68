// Timestep embedding (sinusoid -> MLP -> all 32 resnet projections) and the head_dim-8 self-attention core.
// Reference semantics: diffusers get_timestep_embedding(flip_sin_to_cos=True, freq_shift=0), TimestepEmbedding,
// ResnetBlock2D.time_emb_proj(silu(emb)) and Attention/AttnProcessor2_0 (oracle/unet_oracle.py restates them;
// reached from audiodiffusion/pipeline_audio_diffusion.py:163).
#include "kernels.cuh"

namespace b200ad {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
  return v;
}

// one CTA per sample: emb[dim0] -> h1[4 dim0] -> temb_act[4 dim0]
__global__ void __launch_bounds__(256) temb_mlp_kernel(const float* __restrict__ t, int dim0, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ temb_act) {
  extern __shared__ float tsm[];
  float* emb = tsm;          // dim0
  float* h1 = tsm + dim0;    // 4*dim0
  const int n = blockIdx.x, D = 4 * dim0, half = dim0 / 2;
  const float tv = t[n];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float freq = expf(-logf(10000.0f) * (float)i / (float)half);
    const float a = tv * freq;
    emb[i] = cosf(a);          // flip_sin_to_cos: [cos | sin]
    emb[half + i] = sinf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < D; r += nw) {
    float s = 0.f;
    for (int k = lane; k < dim0; k += 32) s += w1[(long long)r * dim0 + k] * emb[k];
    s = warp_sum(s);
    if (lane == 0) h1[r] = silu_f(s + b1[r]);
  }
  __syncthreads();
  for (int r = warp; r < D; r += nw) {
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s += w2[(long long)r * D + k] * h1[k];
    s = warp_sum(s);
    if (lane == 0) temb_act[(long long)n * D + r] = silu_f(s + b2[r]);
  }
}

// one warp per projection row, all samples
__global__ void __launch_bounds__(256) temb_proj_kernel(const float* __restrict__ temb_act, int N, int D,
                                                        const float* __restrict__ wcat, const float* __restrict__ bcat,
                                                        int rows, float* __restrict__ proj) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  float wv[32];  // D <= 1024
  const int per = D / 32;
  for (int k = 0; k < per; ++k) wv[k] = wcat[(long long)r * D + k * 32 + lane];
  const float bias = bcat[r];
  for (int n = 0; n < N; ++n) {
    float s = 0.f;
    for (int k = 0; k < per; ++k) s += wv[k] * temb_act[(long long)n * D + k * 32 + lane];
    s = warp_sum(s);
    if (lane == 0) proj[(long long)n * rows + r] = s + bias;
  }
}

cudaError_t launch_temb(const float* t, int N, int dim0, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* temb_act, const float* wcat, const float* bcat, int rows,
                        float* proj, cudaStream_t s) {
  const int D = 4 * dim0;
  if (D > 1024 || (D % 32) != 0) return cudaErrorInvalidValue;
  temb_mlp_kernel<<<N, 256, (dim0 + D) * sizeof(float), s>>>(t, dim0, w1, b1, w2, b2, temb_act);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  temb_proj_kernel<<<(rows + 7) / 8, 256, 0, s>>>(temb_act, N, D, wcat, bcat, rows, proj);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ attention
// grid (heads, N). K and V of one (sample, head) are staged in shared memory as fp32; one thread per query.
__global__ void __launch_bounds__(256) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                        __nv_bfloat16* __restrict__ out, int N, int C, int H, int W) {
  extern __shared__ float asm_[];
  const int seq = H * W;
  float* ks = asm_;            // [seq][8]
  float* vs = asm_ + seq * 8;  // [seq][8]
  const Geom g = make_geom(N, H, W);
  const int head = blockIdx.x, n = blockIdx.y;
  const int planes = C >> 3;
  const __nv_bfloat16* base = qkv + (long long)n * 3 * planes * g.PL * 8;
  const __nv_bfloat16* qp = base + (long long)head * g.PL * 8;
  const __nv_bfloat16* kp = base + (long long)(planes + head) * g.PL * 8;
  const __nv_bfloat16* vp = base + (long long)(2 * planes + head) * g.PL * 8;
  for (int p = threadIdx.x; p < seq; p += blockDim.x) {
    const long long pix = (long long)(g.lead + (p / W) * g.Wp + (p % W)) * 8;
    const uint4 kv = *reinterpret_cast<const uint4*>(kp + pix);
    const uint4 vv = *reinterpret_cast<const uint4*>(vp + pix);
    const uint32_t ku[4] = {kv.x, kv.y, kv.z, kv.w}, vu[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = unpack_bf16x2(ku[e]);
      ks[p * 8 + 2 * e] = f.x; ks[p * 8 + 2 * e + 1] = f.y;
      f = unpack_bf16x2(vu[e]);
      vs[p * 8 + 2 * e] = f.x; vs[p * 8 + 2 * e + 1] = f.y;
    }
  }
  __syncthreads();
  const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 8^-0.5 * log2(e)
  for (int p = threadIdx.x; p < seq; p += blockDim.x) {
    const long long pix = (long long)(g.lead + (p / W) * g.Wp + (p % W)) * 8;
    const uint4 qv = *reinterpret_cast<const uint4*>(qp + pix);
    const uint32_t qu[4] = {qv.x, qv.y, qv.z, qv.w};
    float q[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_bf16x2(qu[e]);
      q[2 * e] = f.x * sc; q[2 * e + 1] = f.y * sc;
    }
    float mx = -INFINITY;
    for (int j = 0; j < seq; ++j) {
      const float4 k0 = *reinterpret_cast<const float4*>(ks + j * 8), k1 = *reinterpret_cast<const float4*>(ks + j * 8 + 4);
      const float s = q[0] * k0.x + q[1] * k0.y + q[2] * k0.z + q[3] * k0.w + q[4] * k1.x + q[5] * k1.y + q[6] * k1.z + q[7] * k1.w;
      mx = fmaxf(mx, s);
    }
    float den = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < seq; ++j) {
      const float4 k0 = *reinterpret_cast<const float4*>(ks + j * 8), k1 = *reinterpret_cast<const float4*>(ks + j * 8 + 4);
      const float s = q[0] * k0.x + q[1] * k0.y + q[2] * k0.z + q[3] * k0.w + q[4] * k1.x + q[5] * k1.y + q[6] * k1.z + q[7] * k1.w;
      const float pj = exp2f(s - mx);
      den += pj;
      const float4 v0 = *reinterpret_cast<const float4*>(vs + j * 8), v1 = *reinterpret_cast<const float4*>(vs + j * 8 + 4);
      o[0] += pj * v0.x; o[1] += pj * v0.y; o[2] += pj * v0.z; o[3] += pj * v0.w;
      o[4] += pj * v1.x; o[5] += pj * v1.y; o[6] += pj * v1.z; o[7] += pj * v1.w;
    }
    const float inv = 1.0f / den;
    uint4 ov;
    ov.x = pack_bf16x2(o[0] * inv, o[1] * inv); ov.y = pack_bf16x2(o[2] * inv, o[3] * inv);
    ov.z = pack_bf16x2(o[4] * inv, o[5] * inv); ov.w = pack_bf16x2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(out + ((long long)n * planes + head) * g.PL * 8 + pix) = ov;
  }
}

cudaError_t launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int N, int C, int H, int W, cudaStream_t s) {
  const int seq = H * W;
  const size_t smem = (size_t)seq * 16 * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  dim3 grid(C >> 3, N);
  const int threads = seq >= 256 ? 256 : ((seq + 31) / 32) * 32;
  attention_kernel<<<grid, threads, smem, s>>>(qkv, out, N, C, H, W);
  return cudaGetLastError();
}

}  // namespace b200ad
