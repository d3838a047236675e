// conv_rs — 3x3 stride-1 'same' convolution as a ROW-STREAMING implicit GEMM whose three kernel rows are stacked on
// the UMMA N axis.
//
// Why: with Cout = 48 the tap-by-tap formulation of conv_tc.cu issues 27 MMAs of N = 48 per 128 pixels; each re-reads a
// 4 KB A tile from shared memory for 24 cycles of math, so the tensor pipe waits on shared-memory operand bandwidth
// (ncu: tensor pipe 36 % busy, shared-memory operand reads 60 %).  Here one MMA multiplies a 128-pixel segment of ONE
// INPUT ROW y by the weights of all three kernel rows at once:
//
//   D[128 px x (3 x Cout)] += X[y, x0+dx-1 .. +128, 16 ch] * [ W(kh=0,dx) | W(kh=1,dx) | W(kh=2,dx) ]
//
// The three N blocks are the contributions of input row y to the output rows y+1, y and y-1.  Output-row accumulators
// sit side by side in TMEM in DESCENDING row order, so this 3*Cout-column window is contiguous and simply slides down
// by Cout columns per input row (a ring of 512/Cout slots; a window that would wrap, touch rows outside the strip, or mix
// "first write" with "accumulate" is issued as two or three narrower MMAs).  Per input row that is 9 MMAs of N = 144
// instead of 27 of N = 48: A-operand traffic drops 3x, total shared-memory operand traffic ~2x, same math.
//
// Work unit = (image n, 128-pixel column strip cx, row y); CTA b of G owns the contiguous unit range
// [U*b/G, U*(b+1)/G): perfect balance, each input row is fetched from HBM once (+ one halo row per strip end).
//   warp 0      : TMA producer — one 5-D box per input row: the tensor is viewed as [n][plane][H][W/8][8 px x 8 ch], the
//                 box {64, 18, 1, cin/8, 1} lands as [plane][18 groups][128 B]: the 128-pixel segment plus 8 pixels on each
//                 side (halo of 1 needed), 8 pixels x 16 B = one canonical K-major core matrix; the dx tap is a 16-byte
//                 shift of the descriptor start address.  Left/right zero padding = TMA out-of-bounds fill; rows
//                 outside the image are skipped altogether.
//   warp 1      : tcgen05.mma issuer (whole warp walks the loop, one elected lane issues).
//   warp 2      : TMEM allocation.
//   warps 4..19 : four epilogue warpgroups; warpgroup w drains output rows q = w, w+4, ... of this CTA
//                 (tcgen05.ld -> kernels.cuh::epilogue8 -> 16-byte stores).
// Accumulation order per output pixel: input rows y-1, y, y+1, each over (dx, 16-channel step) — identical to
// conv_tc's (dy, dx, k) order and independent of the strip partition, so results do not depend on tile origin.
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kRsWG = 4;                       // epilogue warpgroups
constexpr int kRsThreads = 128 + 128 * kRsWG;  // 640
constexpr int kRsGroups = 18;                  // 8-pixel groups per staged row: 1 left + 16 + 1 right
constexpr uint32_t kRsPlaneBytes = kRsGroups * 128u;
constexpr int kRsMaxSlots = 32;
constexpr uint32_t kRsAlign = 1024;

__host__ __device__ inline uint32_t rs_align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

struct Strip {
  int n, cx, y0, y1;
};
// next maximal run of units inside one (n, cx) column; advances u
__device__ __forceinline__ bool next_strip(const ConvRsParams& p, int& u, int u1, Strip& s) {
  if (u >= u1) return false;
  const int t = u / p.H;
  s.y0 = u - t * p.H;
  s.n = t / p.cols;
  s.cx = t - s.n * p.cols;
  s.y1 = min(p.H, s.y0 + (u1 - u));
  u += s.y1 - s.y0;
  return true;
}

// KS > 0: cin == 16 * KS known at compile time; NCH > 0: npad == 16 * NCH.  With both, steady-state rows are issued
// as straight-line code (the single issuing thread must stay far below ~70 cycles per MMA).
template <int KS, int NCH, int ACT, int COMB, int EXT>
__global__ void __launch_bounds__(kRsThreads, 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ ConvRsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int S = p.stages;
  const int NS = NCH > 0 ? (512 / (16 * (NCH > 0 ? NCH : 1)) > kRsMaxSlots ? kRsMaxSlots : 512 / (16 * (NCH > 0 ? NCH : 1))) : p.nslots;
  const int NP = NCH > 0 ? 16 * NCH : p.np;

  const uint32_t w_al = rs_align_up(p.wbytes, kRsAlign);
  const uint32_t st_al = rs_align_up(p.stage_bytes, kRsAlign);
  uint8_t* const wsm = smem;
  uint8_t* const stage0 = smem + w_al;
  float* const bias_sm = reinterpret_cast<float*>(stage0 + (size_t)S * st_al);
  float* const slope_sm = bias_sm + NP;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(slope_sm + NP);
  uint64_t* const full = bars;
  uint64_t* const empty = bars + S;
  uint64_t* const tfull = bars + 2 * S;
  uint64_t* const tempty = tfull + kRsMaxSlots;
  uint64_t* const wbar = tempty + kRsMaxSlots;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < NS; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);  // the four warps of the warpgroup that drained the slot
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
    // the packed weights are not produced by the previous kernel: fetch them before the grid-dependency wait
    prefetch_tmap(&src_map);
    mbar_expect_tx(wbar, p.wbytes);
    for (uint32_t off = 0; off < p.wbytes; off += 32768u) {
      const uint32_t len = min(32768u, p.wbytes - off);
      bulk_load_1d(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, len, wbar);
    }
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

  pdl_wait();  // everything below reads or writes activation buffers
  const int u0 = (int)((long long)p.units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)p.units * (blockIdx.x + 1) / gridDim.x);

  // Hand-shakes (the tcgen05.mma issuing thread is the scarce resource — ~48 cycles per MMA issue, ~40 per barrier
  // poll, ~68 per tcgen05.commit against 72 cycles of math per N = 144 MMA — so it only polls `full` and commits
  // `tfull`):
  //   full[st]     producer -> MMA     TMA bytes of an input row have landed AND the accumulator slot that row writes
  //                                    first has been drained + cleared (the producer checks tempty before loading)
  //   tfull[slot]  MMA -> epilogue     tcgen05.commit: every contribution to an output row has completed
  //   tempty[slot] epilogue -> producer  slot read out and zeroed
  //   empty[st]    epilogue -> producer  the MMAs that read a stage have completed (implied by the tfull the epilogue saw)
  if (warp == 0) {
    if (lane == 0) {
      int st = 0, qbase = 0, u = u0;
      uint32_t st_par = 1;
      Strip s;
      while (next_strip(p, u, u1, s)) {
        const int ya = max(s.y0 - 1, 0), yb = min(s.y1, p.H - 1);
        for (int yi = ya; yi <= yb; ++yi) {
          mbar_wait_parked(&empty[st], st_par);
          // output rows that receive their first contribution from this input row: yi + 1, and row 0 at yi == 0
          for (int r = (yi == 0 ? 0 : yi + 1); r <= yi + 1; ++r)
            if (r >= s.y0 && r < s.y1) {
              const int q = qbase + (r - s.y0);
              mbar_wait_parked(&tempty[q % NS], (((uint32_t)(q / NS)) & 1u) ^ 1u);
            }
          mbar_expect_tx(&full[st], p.stage_bytes);
          tma_load_5d(stage0 + (size_t)st * st_al, &src_map, &full[st], 0, s.cx * 16 - 1, yi, p.src_plane0, s.n);
          if (++st == S) st = 0, st_par ^= 1u;
        }
        qbase += s.y1 - s.y0;
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint32_t idesc0 = make_idesc_bf16(128, 0);
    const uint32_t b_lbo = (uint32_t)(3 * NP) * 16u;  // next 8-input-channel slab
    const uint64_t db0 = make_smem_desc(smem_u32(wsm), b_lbo, 128u);
    const uint32_t b_lo0 = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    const uint64_t da0 = make_smem_desc(smem_u32(stage0) + 7u * 16u, kRsPlaneBytes, 128u);
    const uint32_t a_lo0 = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32);
    const uint32_t st_units = st_al >> 4;
    const int ksteps = KS > 0 ? KS : (p.cin >> 4);
    const int cin8 = 2 * ksteps;
    const uint32_t tfull_s = smem_u32(tfull);
    int st = 0, qbase = 0, u = u0;
    uint32_t st_par = 0;
    Strip s;
    while (next_strip(p, u, u1, s)) {
      const int ya = max(s.y0 - 1, 0), yb = min(s.y1, p.H - 1);
      for (int yi = ya; yi <= yb; ++yi) {
        const uint32_t a_lo = a_lo0 + (uint32_t)st * st_units;
        const int qn = qbase + (yi + 1 - s.y0);  // CTA-local index of output row yi + 1
        const int slot_n = qn % NS;
        mbar_wait(&full[st], st_par);
        tc_fence_after();
        // every accumulator slot is zero when it is handed over (the epilogue clears it after reading): all MMAs accumulate
        if (KS > 0 && NCH > 0 && yi - 1 >= s.y0 && yi + 1 < s.y1 && slot_n >= 2) {
          // steady state: rows yi+1, yi, yi-1 sit in three consecutive slots -> one N = 3 * npad MMA per (dx, k step)
          constexpr uint32_t kN = 16u * (NCH > 0 ? NCH : 1);
          constexpr uint32_t kI3 = make_idesc_bf16(128, (int)(3 * kN));
          const uint32_t d0 = tmem_base + (uint32_t)(NS - 1 - slot_n) * kN;
          if (leader) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int kk = 0; kk < (KS > 0 ? KS : 1); ++kk)
                umma_bf16_lohi<true>(d0, a_lo + (uint32_t)(dx + kk * 2 * (int)(kRsPlaneBytes >> 4)), a_hi,
                                     b_lo0 + (uint32_t)((dx * 2 * KS + 2 * kk) * 3 * (int)kN), b_hi, kI3);
            umma_commit_addr(tfull_s + 8u * (uint32_t)(slot_n - 2));  // output row yi - 1 is complete
          }
        } else {
          // strip / image borders and ring wrap: N block g (kernel row kh = g) feeds output row r = yi + 1 - g
          bool v[3];
          uint32_t col[3];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int r = yi + 1 - g;
            v[g] = r >= s.y0 && r < s.y1;
            const int slot = v[g] ? (qn - g) % NS : 0;
            col[g] = (uint32_t)((NS - 1 - slot) * NP);
          }
          // neighbouring blocks whose slots are contiguous go out as one MMA
          const bool m01 = v[0] && v[1] && col[1] == col[0] + (uint32_t)NP;
          const bool m12 = v[1] && v[2] && col[2] == col[1] + (uint32_t)NP;
          const int n0 = v[0] ? 1 + (m01 ? 1 + (m12 ? 1 : 0) : 0) : 0;
          const int n1 = (v[1] && !m01) ? 1 + (m12 ? 1 : 0) : 0;
          const int n2 = (v[2] && !m12) ? 1 : 0;
          uint32_t idg[3];
          idg[0] = n0 ? idesc0 | ((uint32_t)((n0 * NP) >> 3) << 17) : 0u;
          idg[1] = n1 ? idesc0 | ((uint32_t)((n1 * NP) >> 3) << 17) : 0u;
          idg[2] = n2 ? idesc0 | ((uint32_t)((n2 * NP) >> 3) << 17) : 0u;
          if (leader) {
            uint32_t b_dx = b_lo0;
            for (int dx = 0; dx < 3; ++dx) {
              uint32_t a = a_lo + (uint32_t)dx, b = b_dx;
              for (int kk = 0; kk < ksteps; ++kk) {
#pragma unroll
                for (int g = 0; g < 3; ++g)
                  if (idg[g] != 0u) umma_bf16_lohi<true>(tmem_base + col[g], a, a_hi, b + (uint32_t)(g * NP), b_hi, idg[g]);
                a += 2u * (kRsPlaneBytes >> 4);
                b += 2u * (uint32_t)(3 * NP);
              }
              b_dx += (uint32_t)(cin8 * 3 * NP);
            }
            // output rows whose last contribution this was: r = yi - 1 always, r = yi on the image's last row
            if (v[2]) umma_commit_addr(tfull_s + 8u * (uint32_t)((qn - 2) % NS));
            if (v[1] && yi == p.H - 1) umma_commit_addr(tfull_s + 8u * (uint32_t)((qn - 1) % NS));
          }
        }
        if (++st == S) st = 0, st_par ^= 1u;
      }
      qbase += s.y1 - s.y0;
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 2;
    const int qd = warp & 3;  // TMEM lane quarter this warp may read
    const int cstore = (p.epi.cout + 7) & ~7;
    int qbase = 0, jbase = 0, u = u0;
    Strip s;
    while (next_strip(p, u, u1, s)) {
      const int ya = max(s.y0 - 1, 0), yb = min(s.y1, p.H - 1);
      const int n = s.n;
      const int x = s.cx * 128 + qd * 32 + lane;
      const bool valid = x < p.W;
      for (int y = s.y0 + ((wg - qbase) & (kRsWG - 1)); y < s.y1; y += kRsWG) {
        const int q = qbase + (y - s.y0);  // q % kRsWG == wg
        const int slot = q % NS;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((NS - 1 - slot) * NP);
        const uint32_t par = ((uint32_t)(q / NS)) & 1u;
        constexpr bool kUsesRes = NCH > 0 && (COMB == RSB_COMB_SPAB_GATE || COMB == RSB_COMB_MUL || COMB == RSB_COMB_AXPY);
        uint4 pre[kUsesRes ? 2 * NCH : 1];
        if constexpr (kUsesRes) {
          // the residual does not depend on this row's MMAs: fetch it while they are still running
          if (valid) {
            const T* rp = reinterpret_cast<const T*>(p.epi.res1) + planar_index(n, p.epi.res1_planes, p.epi.res1_plane0, p.H, p.W, y, x);
            const size_t plane_stride = (size_t)p.H * p.W * 8;
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k)
              if (k * 8 < cstore) pre[k] = *reinterpret_cast<const uint4*>(rp + k * plane_stride);
          }
        }
        mbar_wait_parked(&tfull[slot], par);
        tc_fence_after();
        if (qd == 0 && lane == 0) {
          // every MMA up to input row min(y + 1, H - 1) has completed: hand those rows' stages back to the producer
          // (row y + 1 by its predecessor's epilogue; the strip's first output row also covers the rows before it)
          const int lo = y == s.y0 ? ya : y + 1;
          const int hi = min(y + 1, yb);
          for (int yi = lo; yi <= hi; ++yi) mbar_arrive(&empty[(jbase + (yi - ya)) % S]);
        }
        if constexpr (NCH > 0) {
          uint32_t r[2][16];
          tmem_ld16(taddr, r[0]);
          const bool lean = EXT == 0 && p.epi.simple;
          T* const drow = reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0, p.H, p.W, y, x);
          const size_t dstride = (size_t)p.H * p.W * 8;
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            const int c = ci * 16;
            tmem_ld_wait();
            if (ci + 1 < NCH) tmem_ld16(taddr + (uint32_t)(c + 16), r[(ci + 1) & 1]);
            tmem_st16_zero(taddr + (uint32_t)c);  // chunk c is in registers: hand the slot back cleared
            if (ci + 1 == NCH) {
              // the accumulator is free as soon as it has been read and cleared — before this row's math and stores
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[slot]);
            }
            if (lean) {
              if (valid) epilogue16_planar<ACT, COMB>(p.epi, bias_sm, slope_sm, r[ci & 1], c, cstore, drow, dstride, kUsesRes ? &pre[2 * ci] : nullptr);
            } else if (valid) {
              float vv[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[ci & 1][k]);
              if (c < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, c, n, y, x, kUsesRes ? &pre[2 * ci] : nullptr);
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[ci & 1][8 + k]);
              if (c + 8 < cstore)
                epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, c + 8, n, y, x, kUsesRes ? &pre[2 * ci + 1] : nullptr);
            }
          }
        } else {
          for (int c = 0; c < NP; c += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)c, r);
            tmem_ld_wait();
            tmem_st16_zero(taddr + (uint32_t)c);
            if (valid) {
              float vv[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[k]);
              if (c < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, c, n, y, x);
#pragma unroll
              for (int k = 0; k < 8; ++k) vv[k] = __uint_as_float(r[8 + k]);
              if (c + 8 < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_sm, slope_sm, vv, c + 8, n, y, x);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[slot]);
        }
      }
      qbase += s.y1 - s.y0;
      jbase += yb - ya + 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef void (*RsKernelFn)(const CUtensorMap, const ConvRsParams);

struct RsVariant {
  int ks, nch, act, comb, ext;
  RsKernelFn fn;
};

#define RSB_V(KS, NCH, ACT, COMB) {KS, NCH, ACT, COMB, 0, conv_rs_kernel<KS, NCH, ACT, COMB, 0>}
#define RSB_X(KS, NCH, ACT, COMB) {KS, NCH, ACT, COMB, kRuntime, conv_rs_kernel<KS, NCH, ACT, COMB, kRuntime>}
const RsVariant kRsVariants[] = {
    // SPAN / SPANPlus (48 -> 48)
    RSB_V(3, 3, RSB_ACT_SILU, RSB_COMB_NONE),
    RSB_V(3, 3, RSB_ACT_MISH, RSB_COMB_NONE),
    RSB_V(3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    RSB_V(3, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    // first layers on the 16-channel planar copy of the caller's input (3 real channels)
    RSB_V(1, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(1, 4, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(1, 4, RSB_ACT_PRELU, RSB_COMB_NONE),
    // Compact (64 -> 64, PReLU)
    RSB_V(4, 4, RSB_ACT_PRELU, RSB_COMB_NONE),
    // ESRGAN dense blocks (64 + 32k -> 32), trunk / HR convs (64 -> 64), RealPLKSR
    RSB_V(4, 2, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(6, 2, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(8, 2, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(10, 2, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(4, 4, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(4, 4, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(4, 4, RSB_ACT_NONE, RSB_COMB_AXPY),
    RSB_V(6, 4, RSB_ACT_NONE, RSB_COMB_AXPY),  // the two K halves of ESRGAN's 192 -> 64 conv
    RSB_V(4, 4, RSB_ACT_SIGMOID, RSB_COMB_MUL),
    RSB_V(8, 4, RSB_ACT_NONE, RSB_COMB_NONE),
    // upsampler convs storing PixelShuffle'd into the caller's tensor
    RSB_X(3, 1, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_X(4, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    // SPAN / SPANPlus / SpanPP: conv_cat (1x1 over the 4 x 48-channel concat) merged into the upsampler conv: 192 -> 12 / 48
    RSB_X(12, 1, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_X(12, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    // RTMoSR (dim 32, hidden 64): stem on the 16-channel planar copy, the three fc1 parts (the gate part ends in mish(g) * cat(i, c)),
    // fc2 with mish(.) + shortcut, and the same for dim 64 / hidden 128
    RSB_V(1, 2, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(2, 2, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(2, 4, RSB_ACT_MISH, RSB_COMB_MUL),
    RSB_V(4, 2, RSB_ACT_MISH, RSB_COMB_AXPY),
    RSB_V(8, 4, RSB_ACT_MISH, RSB_COMB_AXPY),
    RSB_X(2, 1, RSB_ACT_NONE, RSB_COMB_NONE),
    // GateRV3 (dim 32): SPABs at 32 channels, the grouped 3x3 conv as block-diagonal 32 / 64-channel convs whose first half ends in
    // SimpleGate (x1 * x2)
    RSB_V(2, 2, RSB_ACT_SILU, RSB_COMB_NONE),
    RSB_V(2, 2, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    RSB_V(2, 2, RSB_ACT_NONE, RSB_COMB_MUL),
    RSB_V(4, 4, RSB_ACT_NONE, RSB_COMB_MUL),
    // runtime geometry, specialised epilogue
    RSB_V(0, 0, RSB_ACT_NONE, RSB_COMB_MUL),
    RSB_V(0, 0, RSB_ACT_SILU, RSB_COMB_NONE),
    RSB_V(0, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(0, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_X(0, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    // everything else
    RSB_X(0, 0, kRuntime, kRuntime),
};
#undef RSB_V
#undef RSB_X
constexpr int kNumRsVariants = sizeof(kRsVariants) / sizeof(kRsVariants[0]);

RsKernelFn rs_pick(const ConvRsParams& p) {
  int act = p.epi.act;
  const int comb = p.epi.combine;
  if (comb == RSB_COMB_SPAB_GATE) act = RSB_ACT_NONE;  // the gate ignores `act`
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < kNumRsVariants; ++i) {
      const RsVariant& v = kRsVariants[i];
      const bool geo = pass == 0 ? (v.ks * 16 == p.cin && v.nch * 16 == p.np) : (v.ks == 0 && v.nch == 0);
      const bool ext = p.epi.dst_external ? v.ext == kRuntime : v.ext == 0;
      if (geo && ext && v.act == act && v.comb == comb) return v.fn;
    }
  return kRsVariants[kNumRsVariants - 1].fn;
}

}  // namespace

size_t conv_rs_smem_bytes(int cin, int np, int stages) {
  const uint32_t wbytes = 9u * (uint32_t)cin * (uint32_t)np * 2u;
  const uint32_t stage = (uint32_t)(cin / 8) * kRsPlaneBytes;
  return (size_t)rs_align_up(wbytes, kRsAlign) + (size_t)stages * rs_align_up(stage, kRsAlign) + 2 * np * sizeof(float) +
         (2 * stages + 2 * kRsMaxSlots + 1) * 8 + 16;
}

uint32_t conv_rs_stage_bytes(int cin) { return (uint32_t)(cin / 8) * kRsPlaneBytes; }

cudaError_t conv_rs_configure(size_t max_smem) {
  for (int i = 0; i < kNumRsVariants; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kRsVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_rs(const CUtensorMap& src_map, const ConvRsParams& p, int num_sms, cudaStream_t stream) {
  // always ask for more than half an SM's shared memory: one CTA per SM, so the 512-column TMEM allocation never contends
  size_t smem = conv_rs_smem_bytes(p.cin, p.np, p.stages);
  if (smem < 120 * 1024) smem = 120 * 1024;
  const int grid = p.units < num_sms ? p.units : num_sms;
  RsKernelFn fn = rs_pick(p);
  return launch_pdl(fn, dim3(grid), dim3(kRsThreads), smem, stream, src_map, p);
}

}  // namespace rsb
