// Micro-benchmark replaying the MMA stream of conv_rs (9 MMAs of N=144 per input row) with one factor varied at a time.
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace rsb::ptx;

__global__ void __launch_bounds__(128, 1) bench(int mode, int rows, int stages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, 144);
    const uint32_t wbytes = 41984, st_bytes = 14336;
    const uint64_t db = make_smem_desc(smem_u32(smem), 144 * 16, 128);
    const uint64_t da = make_smem_desc(smem_u32(smem + wbytes) + 7 * 16, 2304, 128);
    long long t0 = 0, t1 = 0;
    if (leader) {
      t0 = clock64();
      int st = 0;
      for (int r = 0; r < rows; ++r) {
        const uint32_t d = tm + ((mode & 4) ? (uint32_t)((7 - (r % 8)) * 48) : 0u);
        const uint64_t a_row = da + (uint64_t)(((mode & 8) ? st : 0) * (st_bytes >> 4));
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const uint64_t a = a_row + (uint64_t)((mode & 1) ? (dx + kk * 288) : dx);
            const uint64_t b = db + (uint64_t)((mode & 2) ? (dx * 6 + 2 * kk) * 144 : kk * 288);
            umma_bf16(d, a, b, idesc, 1u);
          }
        if ((mode & 16) && (r & 1)) { umma_commit(&dummy[0]); umma_commit(&dummy[1]); umma_commit(&dummy[2]); umma_commit(&dummy[3]); }
        if (++st == stages) st = 0;
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
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rows = 1000;
  for (int mode : {0, 1, 2, 3, 4, 7, 8, 15, 16, 31}) {
    bench<<<148, 128, 200 * 1024>>>(mode, rows, 8, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long cyc;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("mode %2d (1 real A offsets, 2 real B offsets, 4 sliding D, 8 rotating stages, 16 commits/2 rows): %7.1f cycles per row (ideal 648)\n", mode, (double)cyc / rows);
  }
  return 0;
}
