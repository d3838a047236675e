// conv_tc2 — the same implicit-GEMM conv as conv_tc.cu, but on CTA pairs: tcgen05.mma.cta_group::2, M = 256.
//
// Two CTAs of one cluster (= the two SMs of a TPC) each stage their own 16x8-pixel halo tile (their 128 rows of A)
// and HALF of the layer's weights (their N/2 columns of B); the leader CTA issues one MMA for both tiles and the
// hardware reads the other half of B from the peer's shared memory.  For small-N layers (SPAN: N = 48) the kernel is
// bound by shared-memory operand bandwidth, so halving the B read (1.5 KB -> 0.75 KB next to the 4 KB A tile per MMA)
// and halving the issue work per SM is a direct gain; it also halves the weight footprint per SM.
//   * full[s]    lives in the leader: both CTAs' TMA loads complete_tx on it (cp.async.bulk.tensor ... cta_group::2)
//   * empty[s], tfull[a] live in both CTAs: the leader's tcgen05.commit is multicast to the pair
//   * tempty[a]  lives in the leader: the epilogue warps of both CTAs arrive on it (remote arrive through mapa)
// Everything else (planar-8 layout, tap addressing, epilogues) is shared with conv_tc.cu.
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {
namespace {

constexpr int kAcc = 4;
constexpr int kThreads2 = 128 + 128 * kAcc;
constexpr uint32_t kAlign = 1024;
__host__ __device__ inline uint32_t align_up2(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: lanes of the single-lane role warps reach this point at different times
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {  // arrive on the same barrier in CTA `cta`
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(ptx::smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's mbarrier (peer bit of the barrier address cleared)
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint64_t* bar) {  // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int KH, int KW, int KSTEPS, int NCH, int ACT, int COMB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  constexpr int NPAD = NCH * 16, NH = NPAD / 2;
  constexpr int HT = kTileH + KH - 1, WT = kTileW + KW - 1;
  constexpr uint32_t kStage = (uint32_t)HT * WT * KSTEPS * 16 * 2;
  constexpr uint32_t kWHalf = (uint32_t)KH * KW * KSTEPS * 2 * NH * 16;  // bytes of this CTA's half of the weights
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int S = p.stages;

  const uint32_t w_al = align_up2(kWHalf, kAlign), st_al = align_up2(kStage, kAlign);
  uint8_t* const wsm = smem;
  uint8_t* const stage0 = smem + w_al;
  float* const bias_sm = reinterpret_cast<float*>(stage0 + (size_t)S * st_al);
  float* const slope_sm = bias_sm + NPAD;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(slope_sm + NPAD);
  uint64_t* const full = bars;  // used in the leader only
  uint64_t* const empty = bars + S;
  uint64_t* const tfull = bars + 2 * S;
  uint64_t* const tempty = tfull + kAcc;  // used in the leader only
  uint64_t* const wbar = tempty + kAcc;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < kAcc; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);  // 4 epilogue warps in each of the two CTAs
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, p.tmem_cols);
    tmem_relinquish2();
  }
  for (int i = threadIdx.x; i < NPAD; i += blockDim.x) {
    bias_sm[i] = p.epi.bias[i];
    slope_sm[i] = p.epi.slopes != nullptr ? p.epi.slopes[i] : 0.0f;
  }
  if (threadIdx.x == 0) {
    mbar_expect_tx(wbar, kWHalf);
    bulk_load_1d(wsm, reinterpret_cast<const uint8_t*>(p.wpack2) + (size_t)rank * kWHalf, kWHalf, wbar);
  }
  mbar_wait(wbar, 0);  // this CTA's weight half has landed
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers initialised and weights resident in BOTH CTAs before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int iters = (p.num_tiles + 2 * npairs - 1 - 2 * pair) / (2 * npairs);  // pair-iterations with a valid rank-0 tile

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&src_map);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < iters; ++j) {
        int tile = 2 * (pair + j * npairs) + (int)rank;
        if (tile >= p.num_tiles) tile = p.num_tiles - 1;  // odd tail: the peer recomputes the last tile and drops it
        mbar_wait(&empty[s], ph ^ 1);
        const bool no_peer = (p.dbg & 16) != 0;
        if (p.dbg & 1024) {  // no TMA at all: plain arrive
          if (rank == 0) mbar_arrive(&full[s]);
          if (++s == S) s = 0, ph ^= 1;
          continue;
        }
        if (rank == 0) mbar_expect_tx(&full[s], no_peer ? kStage : 2 * kStage);
        if (no_peer && rank == 1) { if (++s == S) s = 0, ph ^= 1; continue; }
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        if (p.dbg & 32)
          tma_load_4d(stage0 + (size_t)s * st_al, &src_map, &full[s], 8 * (tx * kTileW - p.pad_l), ty * kTileH - p.pad_t, p.src_plane0, n);
        else
          tma_load_4d_pair(stage0 + (size_t)s * st_al, &src_map, &full[s], 8 * (tx * kTileW - p.pad_l), ty * kTileH - p.pad_t,
                           p.src_plane0, n);
        if (++s == S) s = 0, ph ^= 1;
      }
    }
  } else if ((warp == 1 || warp == 3) && rank == 0) {
    const bool one_warp = (p.dbg & 1) != 0;
    const int first = warp == 1 ? 0 : 1;
    const int step = one_warp ? 1 : 2;
    const bool leader = elect_one() && !(one_warp && warp == 3);
    constexpr uint32_t idesc = make_idesc_bf16(256, NPAD);
    constexpr uint32_t kPlane = (uint32_t)(HT * WT);
    const uint64_t db = make_smem_desc(smem_u32(wsm), (uint32_t)NH * 16u, 128u);
    constexpr uint32_t b_kstep = 2u * (uint32_t)NH;  // descriptor units (16 B) per 16-channel K step
    for (int i = one_warp ? 0 : first; i < (one_warp && warp == 3 ? 0 : iters); i += step) {
      const int s = i % S, acc = i % kAcc;
      if (p.dbg & 256) __nanosleep(20000);
      mbar_wait(&tempty[acc], (((uint32_t)(i / kAcc)) & 1u) ^ 1u);
      mbar_wait(&full[s], (uint32_t)(i / S) & 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)acc * NPAD;
      const uint64_t da = make_smem_desc(smem_u32(stage0 + (size_t)s * st_al), kPlane * 16u, (uint32_t)WT * 16u);
      if (leader) {
#pragma unroll
        for (int dy = 0; dy < KH; ++dy)
#pragma unroll
          for (int dx = 0; dx < KW; ++dx)
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
              const uint64_t a = da + (uint64_t)((uint32_t)(dy * WT + dx) + 2u * kk * kPlane);
              const uint64_t b = db + (uint64_t)((uint32_t)((dy * KW + dx) * KSTEPS + kk) * b_kstep);
              if (!(p.dbg & 8)) umma2_bf16(d, a, b, idesc, (dy | dx | kk) != 0 ? 1u : 0u);
            }
        if (p.dbg & 128) {  // software multicast: plain arrives on both CTAs' barriers
          mbar_arrive_cta(&empty[s], 0), mbar_arrive_cta(&empty[s], 1);
          mbar_arrive_cta(&tfull[acc], 0), mbar_arrive_cta(&tfull[acc], 1);
        } else {
          umma2_commit_pair(&empty[s]);
          umma2_commit_pair(&tfull[acc]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2, q = warp & 3;
    const int row = q * 32 + lane, ry = row >> 3, rx = row & 7;
    const int cstore = (p.epi.cout + 7) & ~7;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * NPAD;
    constexpr bool kUsesRes = COMB == RSB_COMB_SPAB_GATE || COMB == RSB_COMB_MUL || COMB == RSB_COMB_AXPY;
    uint32_t aph = 0;
    for (int j = g; j < iters; j += kAcc) {
      const int tile = 2 * (pair + j * npairs) + (int)rank;
      const bool tile_ok = tile < p.num_tiles;
      const int tl = tile_ok ? tile : p.num_tiles - 1;
      const int n = tl / tiles_per_img;
      const int rem = tl - n * tiles_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int y = ty * kTileH + ry, x = tx * kTileW + rx;
      const bool valid = tile_ok && (y < p.H) && (x < p.W) && !(p.dbg & 4);
      uint4 pre[kUsesRes ? 2 * NCH : 1];
      if constexpr (kUsesRes) {
        if (valid) {
          const T* rp = reinterpret_cast<const T*>(p.epi.res1) + planar_index(n, p.epi.res1_planes, p.epi.res1_plane0, p.H, p.W, y, x);
          const size_t plane_stride = (size_t)p.H * p.W * 8;
#pragma unroll
          for (int c = 0; c < 2 * NCH; ++c)
            if (c * 8 < cstore) pre[c] = *reinterpret_cast<const uint4*>(rp + c * plane_stride);
        }
      }
      mbar_wait(&tfull[g], aph);
      tc_fence_after();
      uint32_t r[2][16];
      if (!(p.dbg & 64)) tmem_ld16(taddr, r[0]);
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        if (p.dbg & 64) break;
        const int c = ci * 16;
        tmem_ld_wait();
        if (ci + 1 < NCH) tmem_ld16(taddr + (uint32_t)(c + 16), r[(ci + 1) & 1]);
        if (valid) {
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[ci & 1][k]);
          if (c < cstore) epilogue8<T, true, ACT, COMB, 0>(p.epi, bias_sm, slope_sm, v, c, n, y, x, kUsesRes ? &pre[2 * ci] : nullptr);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[ci & 1][8 + k]);
          if (c + 8 < cstore) epilogue8<T, true, ACT, COMB, 0>(p.epi, bias_sm, slope_sm, v, c + 8, n, y, x, kUsesRes ? &pre[2 * ci + 1] : nullptr);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty[g], 0);  // the leader's MMA warps own the accumulator hand-back
      aph ^= 1;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal into it
  if (warp == 2) tmem_dealloc2(tmem_base, p.tmem_cols);
}

typedef void (*KernelFn2)(const CUtensorMap, const ConvTcParams);
struct Variant2 {
  int kh, kw, ksteps, nch, act, comb;
  KernelFn2 fn;
};
#define RSB_V2(KH, KW, KS, NCH, ACT, COMB) {KH, KW, KS, NCH, ACT, COMB, conv_tc2_kernel<KH, KW, KS, NCH, ACT, COMB>}
const Variant2 kVariants2[] = {
    RSB_V2(3, 3, 3, 3, RSB_ACT_SILU, RSB_COMB_NONE),
    RSB_V2(3, 3, 3, 3, RSB_ACT_MISH, RSB_COMB_NONE),
    RSB_V2(3, 3, 3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    RSB_V2(3, 3, 3, 3, RSB_ACT_NONE, RSB_COMB_NONE),
};
#undef RSB_V2
constexpr int kNumVariants2 = sizeof(kVariants2) / sizeof(kVariants2[0]);

size_t smem_bytes2(int kh, int kw, int ksteps, int npad, int stages) {
  const uint32_t whalf = (uint32_t)kh * kw * ksteps * 2 * (npad / 2) * 16;
  const uint32_t stage = (uint32_t)(kTileH + kh - 1) * (kTileW + kw - 1) * ksteps * 16 * 2;
  return (size_t)align_up2(whalf, kAlign) + (size_t)stages * align_up2(stage, kAlign) + 2 * npad * sizeof(float) +
         (2 * stages + 2 * kAcc + 1) * 8 + 16;
}

}  // namespace

bool conv_tc2_supported(const ConvTcParams& p) {
  if (p.wpack2 == nullptr || p.epi.dst_external || p.nchunks != 1 || p.num_tiles < 4) return false;
  int act = p.epi.act;
  if (p.epi.combine == RSB_COMB_SPAB_GATE) act = RSB_ACT_NONE;
  for (int i = 0; i < kNumVariants2; ++i) {
    const Variant2& v = kVariants2[i];
    if (v.kh == p.kh && v.kw == p.kw && v.ksteps == (p.cin >> 4) && v.nch * 16 == p.npad && v.act == act && v.comb == p.epi.combine)
      return true;
  }
  return false;
}

cudaError_t conv_tc2_configure(size_t max_smem) {
  for (int i = 0; i < kNumVariants2; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kVariants2[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_tc2(const CUtensorMap& src_map, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  int act = p.epi.act;
  if (p.epi.combine == RSB_COMB_SPAB_GATE) act = RSB_ACT_NONE;
  KernelFn2 fn = nullptr;
  for (int i = 0; i < kNumVariants2; ++i) {
    const Variant2& v = kVariants2[i];
    if (v.kh == p.kh && v.kw == p.kw && v.ksteps == (p.cin >> 4) && v.nch * 16 == p.npad && v.act == act && v.comb == p.epi.combine)
      fn = v.fn;
  }
  if (!fn) return cudaErrorInvalidValue;
  const int grid = (num_sms / 2) * 2;
  const size_t smem = smem_bytes2(p.kh, p.kw, p.cin >> 4, p.npad, p.stages);
  static const int dbg = getenv("RSB_TC2_DBG") ? atoi(getenv("RSB_TC2_DBG")) : 0;
  ConvTcParams q = p;
  q.dbg = dbg;
  if (dbg & 2) q.stages = 2;
  const size_t smem2 = smem_bytes2(q.kh, q.kw, q.cin >> 4, q.npad, q.stages);
  const int grid2 = (dbg & 512) ? grid / 2 : grid;
  fn<<<grid2, kThreads2, smem2, stream>>>(src_map, q);
  return cudaGetLastError();
}

}  // namespace rsb
