// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode, cta_group::1) as a function of N and of the
// operand layout, one CTA per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../resselt_b200/csrc umma_n.cu -o umma_n
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace rsb::ptx;

__global__ void __launch_bounds__(128, 1) bench(int N, int reps, int a_sbo, int a_lbo, int a_shift_units, int nacc, long long* out, int ncommit = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t da = make_smem_desc(smem_u32(smem), a_lbo, a_sbo);
    const uint64_t db = make_smem_desc(smem_u32(smem + 24 * 1024), N * 16, 128);
    long long t0 = 0, t1 = 0;
    if (leader) {
      t0 = clock64();
      for (int r = 0; r < reps; r += 9) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const uint32_t d = tm + (uint32_t)(((r + j) % nacc) * N);
          umma_bf16(d, da + (uint64_t)((j % 3) * a_shift_units), db + (uint64_t)((j % 3) * 2 * N), idesc, 1u);
        }
        if (ncommit > 0) umma_commit(&dummy[0]);
        if (ncommit > 1) umma_commit(&dummy[1]);
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (leader) {
      t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 9 * 400;
  for (int nc = 0; nc <= 2; ++nc)
    for (int N : {16, 48, 96, 144, 192, 256}) {
      bench<<<148, 128, 64 * 1024>>>(N, reps, 128, 2304, 1, 1, d, nc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("commits per 9 MMAs=%d N=%3d: %7.1f cycles/MMA  %7.1f cycles per 9-MMA row (ideal %5.1f)\n", nc, N, (double)cyc / reps, 9.0 * cyc / reps, 9 * N / 2.0);
    }
  return 0;
}
