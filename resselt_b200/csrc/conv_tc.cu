// conv_tc — stride-1 'same' convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels x Cout] = sum over taps (dy,dx) and 16-channel K steps of
//                          A_tap[128 pixels x 16 ch] * W_tap[16 ch x Cout]        (tcgen05.mma, kind::f16, bf16 in, fp32 acc)
//
// One CTA per SM, persistent over 16x8-pixel output tiles, warp-specialised:
//   warp 0   : TMA producer. One 4-D box per tile brings the (16+kh-1) x (8+kw-1) halo tile of every input plane
//              into shared memory; out-of-image coordinates are zero-filled by TMA == the conv's zero padding.
//              Because activations are stored planar-8 ([C/8][H][W][8]), the box lands as [plane][row][pixel][8ch]:
//              8 consecutive pixels x 16 B are exactly one canonical no-swizzle K-major core matrix. A tap is then
//              just a different start address inside the same halo tile (+ (dy*WT + dx) * 16 B): the 9 taps of a 3x3
//              re-use one shared-memory copy, no im2col, no per-tap reload.
//   warp 1   : single-thread tcgen05.mma issuer; accumulators live in TMEM (two buffers, so tile i+1's MMAs overlap
//              tile i's epilogue).
//   warp 2   : TMEM allocation / release.
//   warps 4-7: epilogue. tcgen05.ld (32 lanes x 16 columns) -> bias / activation / gate / residual / PixelShuffle
//              (kernels.cuh::epilogue8) -> 16-byte stores, 8 neighbouring pixels filling one 128-byte line.
// The packed weights of the layer ([tap][cin/8][npad][8] bf16, also canonical K-major) stay resident in shared
// memory for the CTA's lifetime (one cp.async.bulk).
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kAlign = 1024;

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

template <typename T, bool kFast>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;

  const uint32_t w_al = align_up(p.wbytes, kAlign);
  const uint32_t st_al = align_up(p.stage_bytes, kAlign);
  uint8_t* const wsm = smem;
  uint8_t* const stage0 = smem + w_al;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)S * st_al);
  uint64_t* const full = bars;
  uint64_t* const empty = bars + S;
  uint64_t* const tfull = bars + 2 * S;
  uint64_t* const tempty = bars + 2 * S + 2;
  uint64_t* const wbar = bars + 2 * S + 4;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4);
    mbar_init(&tempty[1], 4);
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int HT = kTileH + p.kh - 1;
  const int WT = kTileW + p.kw - 1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&src_map);
      mbar_expect_tx(wbar, p.wbytes);
      for (uint32_t off = 0; off < p.wbytes; off += 32768u) {
        const uint32_t len = min(32768u, p.wbytes - off);
        bulk_load_1d(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, len, wbar);
      }
      int i = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], p.stage_bytes);
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        tma_load_4d(stage0 + (size_t)s * st_al, &src_map, &full[s], 8 * (tx * kTileW - p.pad_l),
                    ty * kTileH - p.pad_t, p.src_plane0, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(wbar, 0);
      const uint32_t idesc = make_idesc_bf16(128, p.npad);
      const uint32_t w_base = smem_u32(wsm);
      const uint32_t a_lbo = (uint32_t)(HT * WT) * 16u;  // next 8-channel plane
      const uint32_t a_sbo = (uint32_t)WT * 16u;         // next tile row (8 pixels = one core matrix)
      const uint32_t b_lbo = (uint32_t)p.npad * 16u;     // next 8-input-channel slab
      const uint32_t b_sbo = 128u;                       // next 8 output channels
      const bool swp = p.dbg_swap_lbo_sbo != 0;
      const int ksteps = p.cin >> 4;
      const int cin8 = p.cin >> 3;
      const int taps = p.kh * p.kw;
      int i = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        const int acc = i & 1;
        const uint32_t aph = (i >> 1) & 1;
        mbar_wait(&tempty[acc], aph ^ 1);
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(stage0 + (size_t)s * st_al);
        const uint32_t d = tmem_base + (uint32_t)acc * p.acc_stride;
        uint32_t accum = 0;
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          const uint32_t a_tap = a_base + (uint32_t)(dy * WT + dx) * 16u;
          const uint32_t b_tap = w_base + (uint32_t)(tap * cin8) * b_lbo;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = make_smem_desc(a_tap + (uint32_t)(2 * kk) * a_lbo, swp ? a_sbo : a_lbo, swp ? a_lbo : a_sbo);
            const uint64_t db = make_smem_desc(b_tap + (uint32_t)(2 * kk) * b_lbo, swp ? b_sbo : b_lbo, swp ? b_lbo : b_sbo);
            umma_bf16(d, da, db, idesc, accum);
            accum = 1;
          }
        }
        umma_commit(&empty[s]);    // shared-memory stage may be refilled once these MMAs have read it
        umma_commit(&tfull[acc]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int ry = row >> 3, rx = row & 7;
    const int cstore = (p.epi.cout + 7) & ~7;
    int i = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++i) {
      const int acc = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      mbar_wait(&tfull[acc], aph);
      tc_fence_after();
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y = ty * kTileH + ry, x = tx * kTileW + rx;
      const bool valid = (y < p.H) && (x < p.W);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * p.acc_stride;
      for (int c = 0; c < p.npad; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c, r);
        tmem_ld_wait();
        if (valid) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
          if (c < cstore) epilogue8<T, kFast>(p.epi, v, c, n, y, x);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 + j]);
          if (c + 8 < cstore) epilogue8<T, kFast>(p.epi, v, c + 8, n, y, x);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace

size_t conv_tc_smem_bytes(int cin, int npad, int kh, int kw, int stages) {
  const uint32_t wbytes = (uint32_t)kh * kw * cin * npad * 2u;
  const uint32_t stage = (uint32_t)(kTileH + kh - 1) * (kTileW + kw - 1) * cin * 2u;
  return (size_t)align_up(wbytes, kAlign) + (size_t)stages * align_up(stage, kAlign) + (2 * stages + 5) * 8 + 16;
}

cudaError_t conv_tc_configure(size_t max_smem) {
  return cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)max_smem);
}

cudaError_t launch_conv_tc(const CUtensorMap& src_map, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem = conv_tc_smem_bytes(p.cin, p.npad, p.kh, p.kw, p.stages);
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  static const bool swap_fields = getenv("RSB_DEBUG_DESC_SWAP") != nullptr;
  if (swap_fields) {
    ConvTcParams q = p;
    q.dbg_swap_lbo_sbo = 1;
    conv_tc_kernel<__nv_bfloat16, true><<<grid, kThreads, smem, stream>>>(src_map, q);
    return cudaGetLastError();
  }
  conv_tc_kernel<__nv_bfloat16, true><<<grid, kThreads, smem, stream>>>(src_map, p);
  return cudaGetLastError();
}

}  // namespace rsb
