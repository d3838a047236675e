// How long does the issuing thread spend on k back-to-back tcgen05.mma (pipe drained before)?  Reveals the per-instruction
// issue cost and the depth of the MMA queue.  Also: cost of mbarrier try_wait (already satisfied) and tcgen05.commit.
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace rsb::ptx;

__global__ void __launch_bounds__(128, 1) bench(int N, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[64];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t db = make_smem_desc(smem_u32(smem), N * 16, 128);
    const uint64_t da = make_smem_desc(smem_u32(smem + 32768), 2304, 128);
    if (leader) {
      int nb = 0;
      for (int k = 1; k <= 24; ++k) {
        long long t0 = clock64();
        for (int j = 0; j < k; ++j) umma_bf16(tm, da + (uint64_t)(j % 3), db, idesc, 1u);
        long long t1 = clock64();
        umma_commit(&bar[nb]);
        long long t2 = clock64();
        mbar_wait(&bar[nb], 0);
        long long t3 = clock64();
        ++nb;
        if (blockIdx.x == 0) { out[4 * k] = t1 - t0; out[4 * k + 1] = t2 - t1; out[4 * k + 2] = t3 - t0; }
      }
      // satisfied try_wait cost
      long long t0 = clock64();
      for (int j = 0; j < 16; ++j) mbar_wait(&bar[j], 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      // clock64 overhead
      t0 = clock64();
      long long acc = 0;
      for (int j = 0; j < 16; ++j) acc += clock64();
      t1 = clock64();
      if (blockIdx.x == 0) { out[1] = t1 - t0; out[2] = acc; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 128);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int N : {48, 144, 256}) {
    bench<<<148, 128, 64 * 1024>>>(N, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[128];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("N=%d: 16 satisfied mbar waits %lld cycles; 16 clock64 reads %lld cycles\n", N, h[0], h[1]);
    for (int k = 1; k <= 24; ++k) printf("  k=%2d MMAs: issue %5lld cycles, commit issue %4lld, issue..complete %5lld\n", k, h[4 * k], h[4 * k + 1], h[4 * k + 2]);
  }
  return 0;
}
