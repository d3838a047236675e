// Micro-benchmark: tcgen05.mma (M=128, N=144, K=16, bf16, SS mode, cta_group::1) issued by ONE warp vs TWO warps of the
// same CTA into disjoint TMEM column ranges, one CTA per SM on all SMs.  Question it answers (round 2, fused conv pairs):
// does a second issuing warp add tensor-pipe throughput, or only hide the ~48-cycle issue cost of the first?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../resselt_b200/csrc umma_dual.cu -o umma_dual
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace rsb::ptx;

// mode 0: warp 1 issues `reps` MMAs.  mode 1: warps 1 and 3 issue `reps` MMAs each.  mode 2: warp 1 issues 2 * reps MMAs
// alternating between the two column ranges (what a single-issuer fused pair does).  polls > 0: that many try_wait polls on an
// already-completed barrier before every 9-MMA row (the hand-shake cost of the real kernels).
__global__ void __launch_bounds__(256, 1) bench(int N, int reps, int mode, int polls, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint64_t done_bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 176 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1), mbar_init(&done_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1);
    fence_mbar_init();
    mbar_arrive(&done_bar);  // phase 0 complete: polls with parity 0 succeed at once
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 || (warp == 3 && mode == 1)) {
    const int w = warp == 1 ? 0 : 1;
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N);
    // A: 6 planes x 18 groups x 128 B rows (like the conv kernels); B: [kw][cin/8][N][8]
    const uint64_t da = make_smem_desc(smem_u32(smem + w * 16 * 1024) + 7 * 16, 2304, 128);
    const uint64_t db = make_smem_desc(smem_u32(smem + 32 * 1024 + w * 72 * 1024), N * 16, 128);
    long long t0 = 0;
    unsigned long long g0 = 0;
    if (leader) {
      t0 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      const int total = mode == 2 ? 2 * reps : reps;
      for (int r = 0; r < total; r += 9) {
        for (int k = 0; k < polls; ++k) mbar_wait(&done_bar, 0);
        const int side = mode == 2 ? ((r / 9) & 1) : w;
        const uint32_t d = tm + (uint32_t)(side * 256);
#pragma unroll
        for (int j = 0; j < 9; ++j)
          umma_bf16(d, da + (uint64_t)(j % 3) + (uint64_t)((j / 3) * 2 * (2304 >> 4)), db + (uint64_t)(j * 2 * N), idesc, 1u);
        umma_commit(&dummy[2 * w + ((r / 9) & 1)]);
      }
      umma_commit(&bar[w]);
    }
    mbar_wait(&bar[w], 0);
    if (leader) {
      const long long t1 = clock64();
      unsigned long long g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      if (blockIdx.x == 0) out[2 * w] = t1 - t0, out[2 * w + 1] = (long long)(g1 - g0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 32);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 9 * 2000;
  const char* names[3] = {"one warp, R MMAs", "two warps, R MMAs each", "one warp, 2R MMAs alternating"};
  for (int N : {144, 240})
    for (int polls : {0, 3})
      for (int mode = 0; mode < 3; ++mode) {
        for (int it = 0; it < 2; ++it) {
          bench<<<148, 256, 176 * 1024>>>(N, reps, mode, polls, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[4];
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        if (mode == 1 && h[2] > h[0]) h[0] = h[2], h[1] = h[3];
        const int mmas = mode == 0 ? reps : 2 * reps;
        printf("N=%3d polls=%d %-32s: %8.1f cycles/MMA (ideal %5.1f)  %6.0f MHz  %7.1f TFLOP/s on 148 SMs\n", N, polls, names[mode], (double)h[0] / mmas,
               N / 2.0, 1e3 * h[0] / (double)h[1], 148.0 * mmas * 2.0 * 128 * N * 16 / ((double)h[1] * 1e-9) / 1e12);
      }
  return 0;
}
