// Microbenchmarks that size the conv pipeline: (1) per-SM throughput of 1-D cp.async.bulk (UBLKCP) global->shared
// copies as a function of copy size / alignment / copies in flight, (2) issue-to-retire rate of tcgen05.mma
// (M128 N128 K16, bf16, K-major SWIZZLE_NONE) from fixed shared-memory operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../audio_diffusion_b200/csrc/common.cuh"
using namespace b200ad;

// each CTA: `iters` rounds; per round `ncopies` bulk copies of `bytes` each (issued by `nlanes` lanes), wait, repeat.
__global__ void __launch_bounds__(128, 1) bulk_kernel(const char* src, size_t span, int bytes, int ncopies, int nlanes,
                                                       int misalign, int iters, int depth, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp != 0) return;
  const uint32_t sbase = smem_u32(smem);
  const size_t stage_bytes = (size_t)ncopies * ((bytes + 127) & ~127);
  long long t0 = clock64();
  // `depth` rounds in flight
  for (int it = 0; it < iters + depth; ++it) {
    if (it >= depth) {
      const int s = (it - depth) % depth;
      if (lane == 0) mbar_wait(smem_u32(&bars[s]), ((it - depth) / depth) & 1);
      __syncwarp();
    }
    if (it < iters) {
      const int s = it % depth;
      const uint32_t bar = smem_u32(&bars[s]);
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)bytes * ncopies);
      __syncwarp();
      for (int c = lane; c < ncopies; c += nlanes) {
        if (lane < nlanes) {
          size_t off = ((size_t)(blockIdx.x * 7919 + it * 131 + c * 17) * 4096 + (size_t)misalign) % (span - 65536);
          off = (off & ~(size_t)15);
          if (!misalign) off &= ~(size_t)127;
          bulk_g2s(sbase + s * stage_bytes + c * ((bytes + 127) & ~127), src + off, bytes, bar);
        }
      }
    }
  }
  if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int nmma, int lbo_a, long long* cycles, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 1) tmem_alloc(smem_u32(&tslot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128);
    const uint32_t sb = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < nmma; ++k) {
        const uint64_t a = make_smem_desc(sb + (k % 9) * 16, lbo_a, 128);
        const uint64_t b = make_smem_desc(sb + 32768 + (k % 4) * 4096, 2048, 128);
        umma_bf16(tmem + (k & 3) * 128, a, b, idesc, 1);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), it & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (sink && threadIdx.x == 0) sink[blockIdx.x] = 0.f;
}

int main() {
  const size_t span = (size_t)96 << 20;  // 96 MB source: L2-resident after the first pass
  char* src;
  cudaMalloc(&src, span);
  cudaMemset(src, 1, span);
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  std::vector<long long> h(148);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct Cfg { int bytes, ncopies, nlanes, misalign, depth; };
  const Cfg cfgs[] = {
      {2048, 12, 12, 0, 1},  {2080, 12, 12, 16, 1}, {2080, 12, 1, 16, 1},  {4096, 12, 12, 0, 1},  {36864, 1, 1, 0, 1},
      {24576, 2, 2, 0, 1},   {49152, 1, 1, 0, 1},   {2080, 12, 12, 16, 2}, {2080, 12, 12, 16, 3}, {36864, 1, 1, 0, 2},
      {36864, 1, 1, 0, 3},   {16384, 3, 3, 0, 3},   {8192, 6, 6, 0, 3},    {61440, 1, 1, 0, 3},   {4096, 15, 15, 0, 3},
      {1024, 32, 32, 0, 3},  {2080, 16, 16, 16, 3}, {130 * 16, 6, 6, 16, 3},
  };
  printf("bulk copy: bytes x ncopies (lanes, misalign, depth) -> B/cycle/SM, GB/s chip (at measured clock)\n");
  for (const Cfg& c : cfgs) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      bulk_kernel<<<148, 128, 190 * 1024>>>(src, span, c.bytes, c.ncopies, c.nlanes, c.misalign, iters, c.depth, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += (double)v / 148;
    const double bpc = (double)c.bytes * c.ncopies * iters / avg;
    printf("  %6d x %2d (lanes %2d, mis %2d, depth %d): %7.2f B/cyc/SM   %6.0f cyc/round   ~%5.2f TB/s chip @1.9GHz\n",
           c.bytes, c.ncopies, c.nlanes, c.misalign, c.depth, bpc, avg / iters, bpc * 148 * 1.9e9 / 1e12);
  }
  printf("tcgen05.mma M128 N128 K16 bf16 SWIZZLE_NONE: cycles per MMA (ideal 64)\n");
  for (int lbo : {2080, 2048, 128}) {
    for (int nmma : {4, 16, 36, 144}) {
      const int iters = 500;
      for (int rep = 0; rep < 2; ++rep) {
        mma_kernel<<<148, 128, 64 * 1024>>>(iters, nmma, lbo, cyc, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (auto v : h) avg += (double)v / 148;
      printf("  lbo_a %5d  %3d MMAs per commit: %7.1f cycles per MMA\n", lbo, nmma, avg / iters / nmma);
    }
  }
  return 0;
}
