// Optimizer side of the training step (scripts/train_unet.py:258-267): MSE loss + its gradient, and ONE fused pass over all
// parameters doing clip_grad_norm_(1.0) -> AdamW (decoupled weight decay, bias correction) -> EMAModel.step — instead of
// the ~700-tensor foreach chains torch runs. Two kernels per step: sum of squares of all gradients, then the update.
// Semantics restated in oracle/train_oracle.py (pinned against torch.optim.AdamW there).
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/b200ad.h"
#include "kernels.cuh"

namespace b200ad {
int set_err(const char* fmt, ...);

constexpr int OPT_CHUNK = 16384;   // elements per CTA
constexpr int OPT_THREADS = 256;

struct OptTensor {
  float* p;
  float* m;
  float* v;
  float* ema;       // may be null
  long long n;
};
struct OptChunk {
  int tensor;
  int chunk;        // chunk index inside the tensor
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < OPT_THREADS / 32) t = red[threadIdx.x];
  if (warp == 0) {
#pragma unroll
    for (int sh = 4; sh >= 1; sh >>= 1) t += __shfl_xor_sync(0xffffffffu, t, sh);
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(OPT_THREADS) opt_sqnorm_kernel(const OptTensor* __restrict__ tensors,
                                                                 const OptChunk* __restrict__ chunks,
                                                                 const float* const* __restrict__ grads,
                                                                 double* __restrict__ acc) {
  __shared__ float red[OPT_THREADS / 32];
  const OptChunk c = chunks[blockIdx.x];
  const OptTensor t = tensors[c.tensor];
  const float* g = grads[c.tensor];
  const long long beg = (long long)c.chunk * OPT_CHUNK;
  const long long end = min(t.n, beg + OPT_CHUNK);
  float s = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float x = g[i];
    s = fmaf(x, x, s);
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)tot);
}

struct OptScalars {
  float lr_wd;          // lr * weight_decay
  float beta1, beta2, eps;
  float step_size;      // lr / (1 - beta1^step)
  float inv_sqrt_bc2;   // 1 / sqrt(1 - beta2^step)
  float max_norm;       // <= 0: no clipping
  float one_minus_decay;  // EMA: shadow -= (1 - decay) * (shadow - p); < 0: EMA disabled
};

__global__ void __launch_bounds__(OPT_THREADS) opt_update_kernel(const OptTensor* __restrict__ tensors,
                                                                 const OptChunk* __restrict__ chunks,
                                                                 const float* const* __restrict__ grads,
                                                                 const double* __restrict__ acc, OptScalars hp,
                                                                 float* __restrict__ grad_norm_out) {
  const OptChunk c = chunks[blockIdx.x];
  const OptTensor t = tensors[c.tensor];
  const float* g = grads[c.tensor];
  const float total = (float)sqrt(*acc);
  float coef = 1.f;
  if (hp.max_norm > 0.f) coef = fminf(hp.max_norm / (total + 1e-6f), 1.0f);   // torch.nn.utils.clip_grad_norm_
  if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = total;
  const long long beg = (long long)c.chunk * OPT_CHUNK;
  const long long end = min(t.n, beg + OPT_CHUNK);
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float gi = g[i] * coef;
    float p = t.p[i] * (1.0f - hp.lr_wd);
    const float m = hp.beta1 * t.m[i] + (1.0f - hp.beta1) * gi;
    const float v = hp.beta2 * t.v[i] + (1.0f - hp.beta2) * gi * gi;
    const float denom = sqrtf(v) * hp.inv_sqrt_bc2 + hp.eps;
    p -= hp.step_size * (m / denom);
    t.m[i] = m;
    t.v[i] = v;
    t.p[i] = p;
    if (t.ema && hp.one_minus_decay >= 0.f) {
      const float e = t.ema[i];
      t.ema[i] = e - hp.one_minus_decay * (e - p);
    }
  }
}

// loss = mean((pred - target)^2); grad = 2 (pred - target) / n
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                  size_t n, float inv_n, double* __restrict__ acc, float* __restrict__ grad) {
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - target[i];
    s = fmaf(d, d, s);
    if (grad) grad[i] = 2.0f * inv_n * d;
  }
  const float tot = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)tot);
}
__global__ void mse_final_kernel(const double* acc, float inv_n, float* loss) { *loss = (float)(*acc * (double)inv_n); }

}  // namespace b200ad

using namespace b200ad;

struct b200ad_optim {
  int n_tensors = 0;
  int n_chunks = 0;
  bool has_ema = false;
  OptTensor* d_tensors = nullptr;
  OptChunk* d_chunks = nullptr;
  const float** d_grads = nullptr;
  double* d_acc = nullptr;
};

#define CKT(call)                                                                 \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) return set_err("%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

extern "C" int b200ad_optim_create(int n_tensors, const int64_t* sizes, float* const* params, float* const* exp_avg,
                                   float* const* exp_avg_sq, float* const* ema, b200ad_optim** out) {
  if (n_tensors <= 0 || !sizes || !params || !exp_avg || !exp_avg_sq || !out) return set_err("optim_create: bad argument");
  std::vector<OptTensor> ts(n_tensors);
  std::vector<OptChunk> cs;
  for (int i = 0; i < n_tensors; ++i) {
    ts[i] = OptTensor{params[i], exp_avg[i], exp_avg_sq[i], ema ? ema[i] : nullptr, (long long)sizes[i]};
    const int nc = (int)((sizes[i] + OPT_CHUNK - 1) / OPT_CHUNK);
    for (int c = 0; c < nc; ++c) cs.push_back(OptChunk{i, c});
  }
  b200ad_optim* h = new b200ad_optim();
  h->n_tensors = n_tensors;
  h->n_chunks = (int)cs.size();
  h->has_ema = ema != nullptr;
  CKT(cudaMalloc(&h->d_tensors, ts.size() * sizeof(OptTensor)));
  CKT(cudaMalloc(&h->d_chunks, cs.size() * sizeof(OptChunk)));
  CKT(cudaMalloc(&h->d_grads, (size_t)n_tensors * sizeof(float*)));
  CKT(cudaMalloc(&h->d_acc, 2 * sizeof(double)));
  CKT(cudaMemcpy(h->d_tensors, ts.data(), ts.size() * sizeof(OptTensor), cudaMemcpyHostToDevice));
  CKT(cudaMemcpy(h->d_chunks, cs.data(), cs.size() * sizeof(OptChunk), cudaMemcpyHostToDevice));
  *out = h;
  return 0;
}

extern "C" void b200ad_optim_destroy(b200ad_optim* h) {
  if (!h) return;
  cudaFree(h->d_tensors); cudaFree(h->d_chunks); cudaFree(h->d_grads); cudaFree(h->d_acc);
  delete h;
}

extern "C" int b200ad_optim_step(b200ad_optim* h, const float* const* grads, const b200ad_optim_hparams* hp,
                                 float* grad_norm_out, void* stream) {
  if (!h || !grads || !hp) return set_err("optim_step: null argument");
  if (hp->step < 1) return set_err("optim_step: step is 1-based");
  if (hp->ema_decay >= 0.f && !h->has_ema) return set_err("optim_step: EMA requested but no shadow tensors were bound");
  cudaStream_t st = (cudaStream_t)stream;
  // pageable-host source: the runtime stages the table before returning, so the caller's array may be reused at once
  CKT(cudaMemcpyAsync(h->d_grads, grads, (size_t)h->n_tensors * sizeof(float*), cudaMemcpyHostToDevice, st));
  CKT(cudaMemsetAsync(h->d_acc, 0, sizeof(double), st));
  opt_sqnorm_kernel<<<h->n_chunks, OPT_THREADS, 0, st>>>(h->d_tensors, h->d_chunks, h->d_grads, h->d_acc);
  CKT(cudaGetLastError());
  OptScalars s;
  s.lr_wd = hp->lr * hp->weight_decay;
  s.beta1 = hp->beta1; s.beta2 = hp->beta2; s.eps = hp->eps;
  const double bc1 = 1.0 - pow((double)hp->beta1, (double)hp->step);
  const double bc2 = 1.0 - pow((double)hp->beta2, (double)hp->step);
  s.step_size = (float)((double)hp->lr / bc1);
  s.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  s.max_norm = hp->max_grad_norm;
  s.one_minus_decay = hp->ema_decay >= 0.f ? 1.0f - hp->ema_decay : -1.0f;
  opt_update_kernel<<<h->n_chunks, OPT_THREADS, 0, st>>>(h->d_tensors, h->d_chunks, h->d_grads, h->d_acc, s, grad_norm_out);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int b200ad_mse_loss_grad(const float* pred, const float* target, size_t n, float* loss_out, float* grad_out,
                                    double* scratch, void* stream) {
  if (!pred || !target || !loss_out || !scratch || n == 0) return set_err("mse_loss_grad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CKT(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  const float inv_n = (float)(1.0 / (double)n);
  const int grid = (int)((n + 256 * 8 - 1) / (256 * 8) < 1184 ? (n + 256 * 8 - 1) / (256 * 8) : 1184);
  mse_kernel<<<grid, 256, 0, st>>>(pred, target, n, inv_n, scratch, grad_out);
  CKT(cudaGetLastError());
  mse_final_kernel<<<1, 1, 0, st>>>(scratch, inv_n, loss_out);
  CKT(cudaGetLastError());
  return 0;
}
