// conv_lk — LARGE square kernels (K x K, K odd, 5 <= K <= 17; RealPLKSR's dense 17x17 conv on 16 channels,
// /root/reference/resselt/archs/plksr/rplksr.py:27,36) as a row-streaming implicit GEMM with ALL K kernel rows stacked
// on the UMMA N axis.
//
// Why: with Cout = 16 the tap-by-tap tile kernel issues K*K MMAs of N = 16 per 128 pixels, and a tcgen05.mma occupies the
// tensor pipe for >= ~58 cycles however small its N (tools/ubench/umma_n.cu) — 289 taps x 58 = 16.8k cycles per tile for
// 2.3k cycles of math.  Here one MMA multiplies a 128-pixel segment of ONE INPUT ROW y (shifted by dx) by the kernel
// rows of up to 16 output rows at once:
//
//   D[128 px x (rows x 16)] += X[y, x0 + dx - P .. +128, 16 ch] * [ W(kh = K-1, dx) | ... | W(kh = 0, dx) ]
//
// N block j (kernel row K-1-j) is the contribution of input row y to output row y - P + j.  Output-row accumulators
// (16 fp32 columns each) form a ring of 32 slots in TMEM in ASCENDING row order, so the rows a given input row feeds are
// contiguous columns except where the ring wraps: K = 17 rows go out as two MMAs (9 + 8 rows), three when they wrap.
// Per input row that is K x 2..3 MMAs of N ~ 128..144 instead of K*K of N = 16 per 16 x 8 tile.
//
// The staged input row is the same [plane][18 groups][128 B] box as conv_rs.cu (128 pixels + one 8-pixel group on each
// side): P <= 8 is exactly the guard group, and the dx tap is a 16-byte shift of the descriptor start address.
// Work unit = (image, 128-pixel column strip, output row); a CTA streams a contiguous run of rows and pays 2P extra
// input rows per run.  Accumulation order per output pixel: input rows y-P .. y+P, each over (dx, 16-channel step) —
// the same (kh, kw, k) order as conv_tc.cu, so results do not depend on the kernel that produced them or on tile origin.
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kLkWG = 4;
constexpr int kLkThreads = 128 + 128 * kLkWG;
constexpr int kLkGroups = 18;
constexpr uint32_t kLkPlaneBytes = kLkGroups * 128u;
constexpr int kLkNS = 32;  // accumulator ring: 32 slots x 16 columns = all 512 TMEM columns
constexpr int kLkNP = 16;

__host__ __device__ inline uint32_t lk_align(uint32_t v) { return (v + 1023u) & ~1023u; }

struct LkRun {
  int n, cx, y0, y1;
};
__device__ __forceinline__ bool lk_next_run(const ConvLkParams& p, int& u, int u1, LkRun& s) {
  if (u >= u1) return false;
  const int t = u / p.H;
  s.y0 = u - t * p.H;
  s.n = t / p.cols;
  s.cx = t - s.n * p.cols;
  s.y1 = min(p.H, s.y0 + (u1 - u));
  u += s.y1 - s.y0;
  return true;
}

template <int ACT, int COMB, int EXT>
__global__ void __launch_bounds__(kLkThreads, 1)
conv_lk_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ ConvLkParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  constexpr int NS = kLkNS, NP = kLkNP;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages, K = p.k, P = p.k >> 1;
  pdl_launch_dependents();

  const uint32_t w_al = lk_align(p.wbytes), st_al = lk_align(p.stage_bytes);
  uint8_t* const wsm = smem;
  uint8_t* const stage0 = smem + w_al;
  float* const bias_sm = reinterpret_cast<float*>(stage0 + (size_t)S * st_al);
  float* const slope_sm = bias_sm + NP;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(slope_sm + NP);
  uint64_t* const full = bars;
  uint64_t* const empty = bars + S;
  uint64_t* const tfull = bars + 2 * S;
  uint64_t* const tempty = tfull + NS;
  uint64_t* const wbar = tempty + NS;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < NS; ++a) mbar_init(&tfull[a], 1), mbar_init(&tempty[a], 4);
    mbar_init(wbar, 1);
    fence_mbar_init();
    prefetch_tmap(&src_map);
    mbar_expect_tx(wbar, p.wbytes);
    for (uint32_t off = 0; off < p.wbytes; off += 32768u)
      bulk_load_1d(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, min(32768u, p.wbytes - off), wbar);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    bias_sm[i] = p.epi.bias[i];
    slope_sm[i] = p.epi.slopes != nullptr ? p.epi.slopes[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4 && warp < 8) {
    // all accumulator slots start out zero (afterwards the epilogue clears each slot it has read)
    const uint32_t tz = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < NS * NP; c += 16) tmem_st16_zero(tz + (uint32_t)c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  pdl_wait();
  const int u0 = (int)((long long)p.units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)p.units * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    if (lane == 0) {
      int j = 0, qbase = 0, u = u0;
      LkRun s;
      while (lk_next_run(p, u, u1, s)) {
        const int i0 = max(s.y0 - P, 0), i1 = min(s.y1 + P, p.H);
        for (int yi = i0; yi < i1; ++yi, ++j) {
          const int st = j % S;
          mbar_wait_parked(&empty[st], (((uint32_t)(j / S)) & 1u) ^ 1u);
          // output rows that receive their first contribution from this input row: their slots must have been drained
          const int ra = yi == i0 ? s.y0 : yi + P, rb = min(yi + P, s.y1 - 1);
          for (int r = ra; r <= rb; ++r) {
            const int q = qbase + (r - s.y0);
            mbar_wait_parked(&tempty[q % NS], (((uint32_t)(q / NS)) & 1u) ^ 1u);
          }
          mbar_expect_tx(&full[st], p.stage_bytes);
          tma_load_5d(stage0 + (size_t)st * st_al, &src_map, &full[st], 0, s.cx * 16 - 1, yi, p.src_plane0, s.n);
        }
        qbase += s.y1 - s.y0;
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint32_t idesc0 = make_idesc_bf16(128, 0);
    const int ks = p.cin >> 4;
    const uint32_t slab = (uint32_t)(K * NP);  // 16-byte units per (kw, 8-channel slab): K x 16 weight rows
    const uint64_t db0 = make_smem_desc(smem_u32(wsm), slab * 16u, 128u);
    const uint32_t b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    // dx = 0 reads the pixels x0 - P .. : the staged row starts at x0 - 8
    const uint64_t da0 = make_smem_desc(smem_u32(stage0) + (uint32_t)(8 - P) * 16u, kLkPlaneBytes, 128u);
    const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32);
    const uint32_t st_units = st_al >> 4;
    const uint32_t tfull_s = smem_u32(tfull), empty_s = smem_u32(empty);
    int j = 0, qbase = 0, u = u0;
    LkRun s;
    while (lk_next_run(p, u, u1, s)) {
      const int i0 = max(s.y0 - P, 0), i1 = min(s.y1 + P, p.H);
      for (int yi = i0; yi < i1; ++yi, ++j) {
        const int st = j % S;
        const uint32_t a_lo = a_lo0 + (uint32_t)st * st_units;
        mbar_wait(&full[st], ((uint32_t)(j / S)) & 1u);
        tc_fence_after();
        // output rows fed by this input row, in one or two column-contiguous segments of at most 16 rows
        const int ra = max(yi - P, s.y0), rb = min(yi + P, s.y1 - 1);
        // (a wrap of the ring splits them naturally; 17 unwrapped rows go out as 9 + 8)
        const int wr = rb - ra + 1;
        const int till_wrap = NS - (qbase + (ra - s.y0)) % NS;
        const int c0 = till_wrap < wr ? till_wrap : (wr > 16 ? (wr + 1) / 2 : wr);
        const int nseg = c0 < wr ? 2 : 1;
        const int seg_r[2] = {ra, ra + c0};
        const int seg_n[2] = {c0, wr - c0};
        if (leader) {
          for (int dx = 0; dx < K; ++dx)
            for (int kk = 0; kk < ks; ++kk) {
              const uint32_t a = a_lo + (uint32_t)dx + (uint32_t)kk * 2u * (kLkPlaneBytes >> 4);
              const uint32_t b = b_lo0 + (uint32_t)(dx * 2 * ks + 2 * kk) * slab;
#pragma unroll
              for (int g = 0; g < 2; ++g)
                if (g < nseg) {
                  const int r = seg_r[g];
                  const uint32_t col = (uint32_t)(((qbase + (r - s.y0)) % NS) * NP);
                  umma_bf16_lohi<true>(tmem_base + col, a, a_hi, b + (uint32_t)((r - (yi - P)) * NP), b_hi,
                                       idesc0 | ((uint32_t)((seg_n[g] * NP) >> 3) << 17));
                }
            }
          // the stage is free once these MMAs have completed (a commit per input row is cheap next to its 2K MMAs; releasing
          // stages from the epilogue, as conv_rs does, would need 2P + 1 stages in flight at the start of a run)
          umma_commit_addr(empty_s + 8u * (uint32_t)st);
          // output rows whose last contribution this was: yi - P, and everything still open on the run's last input row
          const int ca = yi - P, cb = yi == i1 - 1 ? s.y1 - 1 : yi - P;
          for (int r = max(ca, s.y0); r <= cb; ++r) umma_commit_addr(tfull_s + 8u * (uint32_t)((qbase + (r - s.y0)) % NS));
        }
      }
      qbase += s.y1 - s.y0;
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int cstore = (p.epi.cout + 7) & ~7;
    int qbase = 0, jbase = 0, u = u0;
    LkRun s;
    while (lk_next_run(p, u, u1, s)) {
      const int i0 = max(s.y0 - P, 0), i1 = min(s.y1 + P, p.H);
      const int n = s.n;
      const int x = s.cx * 128 + qd * 32 + lane;
      const bool valid = x < p.W;
      for (int y = s.y0 + ((wg - qbase) & (kLkWG - 1)); y < s.y1; y += kLkWG) {
        const int q = qbase + (y - s.y0);
        const int slot = q % NS;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(slot * NP);
        mbar_wait_parked(&tfull[slot], ((uint32_t)(q / NS)) & 1u);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        tmem_st16_zero(taddr);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[slot]);
        if (valid) {
          if (EXT == 0 && p.epi.simple && COMB == RSB_COMB_NONE) {
            T* const drow = reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0, p.H, p.W, y, x);
            epilogue16_planar<ACT, COMB>(p.epi, bias_sm, slope_sm, r, 0, cstore, drow, (size_t)p.H * p.W * 8, nullptr);
          } else {
            float vv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[k]);
            epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, 0, n, y, x);
#pragma unroll
            for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[8 + k]);
            if (8 < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, 8, n, y, x);
          }
        }
      }
      qbase += s.y1 - s.y0;
      jbase += i1 - i0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef void (*LkKernelFn)(const CUtensorMap, const ConvLkParams);

LkKernelFn lk_pick(const ConvLkParams& p) {
  if (!p.epi.dst_external && p.epi.act == RSB_ACT_NONE && p.epi.combine == RSB_COMB_NONE) return conv_lk_kernel<RSB_ACT_NONE, RSB_COMB_NONE, 0>;
  return conv_lk_kernel<kRuntime, kRuntime, kRuntime>;
}

}  // namespace

uint32_t conv_lk_weight_bytes(int cin, int k) { return (uint32_t)k * (uint32_t)(cin / 8) * (uint32_t)(k * kLkNP) * 16u; }

size_t conv_lk_smem_bytes(int cin, int k, int stages) {
  const uint32_t stage = (uint32_t)(cin / 8) * kLkPlaneBytes;
  return (size_t)lk_align(conv_lk_weight_bytes(cin, k)) + (size_t)stages * lk_align(stage) + 2 * kLkNP * sizeof(float) +
         (2 * stages + 2 * kLkNS + 1) * 8 + 16;
}

cudaError_t conv_lk_configure(size_t max_smem) {
  cudaError_t e = cudaFuncSetAttribute(conv_lk_kernel<RSB_ACT_NONE, RSB_COMB_NONE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_lk_kernel<kRuntime, kRuntime, kRuntime>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
}

cudaError_t launch_conv_lk(const CUtensorMap& src_map, const ConvLkParams& p, int num_sms, cudaStream_t stream) {
  // more than half an SM's shared memory: one CTA per SM, the 512-column TMEM allocation never contends
  size_t smem = conv_lk_smem_bytes(p.cin, p.k, p.stages);
  if (smem < 120 * 1024) smem = 120 * 1024;
  const int grid = p.units < num_sms ? p.units : num_sms;
  return launch_pdl(lk_pick(p), dim3(grid), dim3(kLkThreads), smem, stream, src_map, p);
}

}  // namespace rsb
