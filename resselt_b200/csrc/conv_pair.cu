// conv_pair — TWO consecutive 3x3 stride-1 'same' convolutions (A, then B) fused into one row-streaming kernel.
//
// Why: an unfused 48->48 3x3 layer at 1080p moves 398 MB through HBM for 86 GFLOP — it sits exactly on the B200 ridge
// (profiles/r1b_rs_bound_switches.md: MMAs alone 63 us, loads + stores alone 76 us, together 79 us).  Here A's activated
// output rows go from TMEM through the epilogue warps straight into a shared-memory ring in the canonical K-major UMMA
// layout and are consumed there by B's MMAs: the intermediate map never touches HBM, the pair is tensor-bound.
//
// Formulation (same as conv_rs.cu): one MMA multiplies a 128-pixel segment of ONE INPUT ROW by the weights of all
// three kernel rows, D[128 px x 3*NP] += X[y] * [W(kh) ...]; the three N blocks are the contributions of input row y
// to output rows y+1, y, y-1.  Differences:
//   * each conv owns exactly THREE TMEM accumulator slots (output row with sequence number q lives in slot q % 3), so
//     the three live rows always fill the conv's 3*NP columns, in one of three cyclic orders.  Instead of splitting
//     the MMA when the order wraps, the weights are stored with five N blocks [W2|W1|W0|W2|W1] (kh index): every
//     cyclic order is a contiguous 3-block window, selected by the start address of the B descriptor.  Every
//     steady-state row is 9 MMAs of N = 3*NP per 16 input channels, no splits.
//   * strips advance by 120 pixels: B's outputs at the first and last pixel of a 128-pixel segment would need A's
//     outputs of the neighbouring strip, so strip cx owns the output pixels [120cx+1, 120cx+121) (image borders are
//     exact: the ring rows keep zero guard pixels = B's zero padding, and A's outputs right of the image are stored as
//     zeros).  Vertically a CTA streams a contiguous run of rows; A runs one row ahead/behind at the run's ends.
//   * per output pixel the accumulation order is input rows y-1, y, y+1, each over (dx, 16-channel step) — identical
//     to conv_rs / conv_tc, and A's output is rounded to bf16 exactly as if it had been stored: the fused pair is
//     bit-identical to the two separate launches.
//
//   warp 0       TMA producer of A's input rows (3-stage ring)
//   warp 1       tcgen05.mma issuer: alternates A(row k) and B(row k - 3)
//   warp 2       TMEM allocation
//   warps 4..11  two warpgroups draining A's accumulators (even / odd rows): bias + activation -> bf16 -> st.shared
//                into the ring -> fence.proxy.async -> mbarrier
//   warps 12..19 two warpgroups draining B's accumulators: the usual fused epilogue (kernels.cuh) -> global
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kPrThreads = 640;
constexpr int kPrGroups = 18;  // 8-pixel groups per staged row: 1 left guard + 16 + 1 right guard
constexpr uint32_t kPrPlaneBytes = kPrGroups * 128u;
constexpr int kPrStages = 3;  // TMA ring of A's input rows
constexpr int kPrDepth = 3;   // ring of A's output rows
constexpr int kPrLag = 3;     // B processes A's output row j at step j + kPrLag
constexpr int kPrStep = 120;  // x distance between strips

__host__ __device__ inline uint32_t pr_align(uint32_t v) { return (v + 127u) & ~127u; }

struct Run {
  int n, cx, y0, y1;
};
__device__ __forceinline__ bool next_run(const ConvPairParams& p, int& u, int u1, Run& s) {
  if (u >= u1) return false;
  const int t = u / p.H;
  s.y0 = u - t * p.H;
  s.n = t / p.cols;
  s.cx = t - s.n * p.cols;
  s.y1 = min(p.H, s.y0 + (u1 - u));
  u += s.y1 - s.y0;
  return true;
}

// MMAs of one input row `yi` of a conv whose output rows [lo, hi) are live; qb = sequence number of output row lo, i0 = the
// run's first input row.
// kZeroed = true  (conv A): every accumulator slot is zero when it is handed over (the epilogue clears it after reading),
//                 so all MMAs accumulate — 9 MMAs of N = 3*NP per row.
// kZeroed = false (conv B): nobody clears accumulators; the first K step of a row's first input row is issued with
//                 accumulate = 0 on that row's N block alone.  That splits one of the nine MMAs in two or three (a
//                 tcgen05.mma occupies the pipe >= ~60 cycles however small its N, tools/ubench/umma_n.cu: ~ +70 cycles
//                 per row), but B's epilogue warps have global loads and stores in flight, and those hold back
//                 tcgen05.wait::st — clearing there put ~1000 cycles on the slot hand-over the MMA thread waits for.
template <int KS, int NP, bool kZeroed>
__device__ __forceinline__ void pair_issue_row(bool leader, uint32_t tmem_c0, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               int yi, int lo, int hi, int qb, int i0) {
  using namespace ptx;
  constexpr uint32_t kPU = kPrPlaneBytes >> 4;  // one 8-channel plane, in 16-byte units
  constexpr uint32_t kI1 = make_idesc_bf16(128, NP);
  if (yi - 1 >= lo && yi + 1 < hi) {
    // rows yi-1, yi, yi+1 fill the three slots; slot order (ascending columns) by s = q(yi) % 3:
    //   s == 1: [yi-1, yi, yi+1] = kh [2,1,0]   s == 0: [yi, yi+1, yi-1] = kh [1,0,2]   s == 2: [yi+1, yi-1, yi] = kh [0,2,1]
    // = weight blocks [rho, rho+1, rho+2] of the five stored blocks [W2|W1|W0|W2|W1]
    const int s = (qb + (yi - lo)) % 3;
    const uint32_t rho = s == 1 ? 0u : (s == 0 ? 1u : 2u);
    constexpr uint32_t kI2 = make_idesc_bf16(128, 2 * NP);
    constexpr uint32_t kI3 = make_idesc_bf16(128, 3 * NP);
    const uint32_t bb = b_lo + rho * (uint32_t)NP;
    if (leader) {
      if (!kZeroed) {
        // first K step: the new output row yi+1 (kernel row 0 = weight block 2) starts from zero
        if (s == 1) {
          umma_bf16_lohi<true>(tmem_c0, a_lo, a_hi, bb, b_hi, kI2);
          umma_bf16_lohi<false>(tmem_c0 + (uint32_t)(2 * NP), a_lo, a_hi, bb + (uint32_t)(2 * NP), b_hi, kI1);
        } else if (s == 2) {
          umma_bf16_lohi<false>(tmem_c0, a_lo, a_hi, bb, b_hi, kI1);
          umma_bf16_lohi<true>(tmem_c0 + (uint32_t)NP, a_lo, a_hi, bb + (uint32_t)NP, b_hi, kI2);
        } else {
          umma_bf16_lohi<true>(tmem_c0, a_lo, a_hi, bb, b_hi, kI1);
          umma_bf16_lohi<false>(tmem_c0 + (uint32_t)NP, a_lo, a_hi, bb + (uint32_t)NP, b_hi, kI1);
          umma_bf16_lohi<true>(tmem_c0 + (uint32_t)(2 * NP), a_lo, a_hi, bb + (uint32_t)(2 * NP), b_hi, kI1);
        }
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
          if (kZeroed || dx + kk > 0)
            umma_bf16_lohi<true>(tmem_c0, a_lo + (uint32_t)dx + (uint32_t)kk * 2u * kPU, a_hi, bb + (uint32_t)((dx * 2 * KS + 2 * kk) * 5 * NP), b_hi, kI3);
    }
  } else {
    // run / image borders: N block of kernel row g feeds output row yi + 1 - g
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int r = yi + 1 - g;
      if (r >= lo && r < hi) {
        const uint32_t col = (uint32_t)(((qb + (r - lo)) % 3) * NP);
        const uint32_t bb = b_lo + (uint32_t)((2 - g) * NP);
        const bool first = !kZeroed && (g == 0 || (g == 1 && yi == i0));  // no earlier input row has contributed to row r
        if (leader) {
          if (first)
            umma_bf16_lohi<false>(tmem_c0 + col, a_lo, a_hi, bb, b_hi, kI1);
          else
            umma_bf16_lohi<true>(tmem_c0 + col, a_lo, a_hi, bb, b_hi, kI1);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
              if (dx + kk > 0)
                umma_bf16_lohi<true>(tmem_c0 + col, a_lo + (uint32_t)dx + (uint32_t)kk * 2u * kPU, a_hi, bb + (uint32_t)((dx * 2 * KS + 2 * kk) * 5 * NP), b_hi, kI1);
        }
      }
    }
  }
}

template <int KS0, int NCH, int ACTM, int ACT, int COMB>
__global__ void __launch_bounds__(kPrThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap res_map, const __grid_constant__ ConvPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  constexpr int NP = 16 * NCH;
  constexpr int S = kPrStages, D = kPrDepth;
  constexpr uint32_t kOrow = (uint32_t)(NP / 8) * kPrPlaneBytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  const uint32_t wA_al = pr_align(p.wbytesA), wB_al = pr_align(p.wbytesB), st_al = pr_align(p.stage_bytes);
  uint8_t* const wA = smem;
  uint8_t* const wB = wA + wA_al;
  uint8_t* const stage0 = wB + wB_al;
  uint8_t* const oring = stage0 + (size_t)S * st_al;
  float* const biasA_sm = reinterpret_cast<float*>(oring + (size_t)D * kOrow);
  float* const biasB_sm = biasA_sm + NP;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(biasB_sm + NP);
  uint64_t* const full = bars;             // [S]   TMA -> MMA
  uint64_t* const empty = full + S;        // [S]   A epilogue -> TMA
  uint64_t* const ofull = empty + S;       // [D]   A epilogue -> MMA   (ring row written)
  uint64_t* const oempty = ofull + D;      // [D]   B epilogue -> A epilogue (ring row consumed)
  uint64_t* const tfullA = oempty + D;     // [3]   MMA -> A epilogue
  uint64_t* const temptyA = tfullA + 3;    // [3]   A epilogue -> MMA
  uint64_t* const tfullB = temptyA + 3;    // [3]
  uint64_t* const temptyB = tfullB + 3;    // [3]
  uint64_t* const wbar = temptyB + 3;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int s = 0; s < D; ++s) mbar_init(&ofull[s], 4), mbar_init(&oempty[s], 1);
    for (int a = 0; a < 3; ++a) {
      mbar_init(&tfullA[a], 1), mbar_init(&temptyA[a], 4);
      mbar_init(&tfullB[a], 1), mbar_init(&temptyB[a], 8);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
    // packed weights are not produced by the previous kernel: fetch them before the grid-dependency wait
    prefetch_tmap(&src_map);
    mbar_expect_tx(wbar, p.wbytesA + p.wbytesB);
    for (uint32_t off = 0; off < p.wbytesA; off += 32768u)
      bulk_load_1d(wA + off, reinterpret_cast<const uint8_t*>(p.wpackA) + off, min(32768u, p.wbytesA - off), wbar);
    for (uint32_t off = 0; off < p.wbytesB; off += 32768u)
      bulk_load_1d(wB + off, reinterpret_cast<const uint8_t*>(p.wpackB) + off, min(32768u, p.wbytesB - off), wbar);
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < NP; i += blockDim.x) {
    biasA_sm[i] = p.biasA[i];
    biasB_sm[i] = p.epi.bias[i];
  }
  // the ring rows' guard groups stay zero for the kernel's lifetime (= zero padding at the image's left / right border)
  for (uint32_t i = threadIdx.x; i < (uint32_t)D * kOrow / 16u; i += blockDim.x) reinterpret_cast<uint4*>(oring)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4 && warp < 8) {
    // conv A's three accumulator slots start out zero (afterwards A's epilogue clears each slot it has read)
    const uint32_t tz = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < 3 * NP; c += 16) tmem_st16_zero(tz + (uint32_t)c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  pdl_wait();  // everything below reads or writes activation buffers
  const int u0 = (int)((long long)p.units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)p.units * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0) {
    if (lane == 0) {
      int j = 0, u = u0;
      Run s;
      while (next_run(p, u, u1, s)) {
        const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
        const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
        // B's residual rows are pulled into L2 several steps before the epilogue warps ask for them (row y is read
        // ~6 steps after A's input row y - 2 is fetched): from HBM they would take longer than the register pipeline
        // of the epilogue can cover
        if (p.res_prefetch)
          for (int yr = s.y0; yr < min(s.y1, iA0 + 2); ++yr) tma_prefetch_l2_5d(&res_map, 0, s.cx * (kPrStep / 8), yr, p.epi.res1_plane0, s.n);
        // ... and so are A's own input rows: with only three shared-memory stages a TMA load has ~2 steps (~2 us) to
        // arrive, which HBM under load does not always meet; from L2 it does
        constexpr int kAhead = 6;
        for (int yp = iA0 + S; yp < min(iA1, iA0 + kAhead); ++yp) tma_prefetch_l2_5d(&src_map, 0, s.cx * (kPrStep / 8) - 1, yp, p.src_plane0, s.n);
        for (int yi = iA0; yi < iA1; ++yi, ++j) {
          if (yi + kAhead < iA1) tma_prefetch_l2_5d(&src_map, 0, s.cx * (kPrStep / 8) - 1, yi + kAhead, p.src_plane0, s.n);
          if (p.res_prefetch && yi + 2 >= s.y0 && yi + 2 < s.y1) tma_prefetch_l2_5d(&res_map, 0, s.cx * (kPrStep / 8), yi + 2, p.epi.res1_plane0, s.n);
          const int st = j % S;
          mbar_wait_parked(&empty[st], (((uint32_t)(j / S)) & 1u) ^ 1u);
          mbar_expect_tx(&full[st], p.stage_bytes);
          tma_load_5d(stage0 + (size_t)st * st_al, &src_map, &full[st], 0, s.cx * (kPrStep / 8) - 1, yi, p.src_plane0, s.n);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint64_t dbA = make_smem_desc(smem_u32(wA), (uint32_t)(5 * NP) * 16u, 128u);
    const uint64_t dbB = make_smem_desc(smem_u32(wB), (uint32_t)(5 * NP) * 16u, 128u);
    const uint64_t daI = make_smem_desc(smem_u32(stage0) + 7u * 16u, kPrPlaneBytes, 128u);
    const uint64_t daO = make_smem_desc(smem_u32(oring) + 7u * 16u, kPrPlaneBytes, 128u);
    const uint32_t bA_lo = (uint32_t)dbA, bA_hi = (uint32_t)(dbA >> 32);
    const uint32_t bB_lo = (uint32_t)dbB, bB_hi = (uint32_t)(dbB >> 32);
    const uint32_t aI_lo = (uint32_t)daI, aI_hi = (uint32_t)(daI >> 32);
    const uint32_t aO_lo = (uint32_t)daO, aO_hi = (uint32_t)(daO >> 32);
    const uint32_t st_units = st_al >> 4, o_units = kOrow >> 4;
    const uint32_t tfullA_s = smem_u32(tfullA), tfullB_s = smem_u32(tfullB);
    int jA = 0, qbA = 0, qbB = 0, u = u0, trow = 0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int y0 = s.y0, y1 = s.y1;
      const int loA = max(y0 - 1, 0), hiA = min(y1 + 1, p.H);
      const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
      const int nAin = iA1 - iA0, nBin = hiA - loA;
      const int lag = p.lag;
      const int steps = max(nAin, nBin + lag);
      for (int k = 0; k < steps; ++k) {
        long long* const tr = (p.trace != nullptr && blockIdx.x == 1 && leader && trow < 128) ? p.trace + 8 * trow : nullptr;
        ++trow;
        if (tr) tr[0] = clock64();
        if (k < nAin) {
          const int yi = iA0 + k;
          // accumulator slots of the output rows whose first contribution comes from this input row
          if (yi == iA0 && yi >= loA) {
            const int q = qbA + (yi - loA);
            mbar_wait(&temptyA[q % 3], (((uint32_t)(q / 3)) & 1u) ^ 1u);
          }
          if (yi + 1 < hiA) {
            const int q = qbA + (yi + 1 - loA);
            mbar_wait(&temptyA[q % 3], (((uint32_t)(q / 3)) & 1u) ^ 1u);
          }
          if (tr) tr[1] = clock64();
          const int st = jA % S;
          mbar_wait(&full[st], ((uint32_t)(jA / S)) & 1u);
          tc_fence_after();
          if (tr) tr[2] = clock64();
          pair_issue_row<KS0, NP, true>(leader, tmem_base, aI_lo + (uint32_t)st * st_units, aI_hi, bA_lo, bA_hi, yi, loA, hiA, qbA, iA0);
          if (leader) {
            if (yi - 1 >= loA) umma_commit_addr(tfullA_s + 8u * (uint32_t)((qbA + (yi - 1 - loA)) % 3));
            if (yi == iA1 - 1 && yi < hiA) umma_commit_addr(tfullA_s + 8u * (uint32_t)((qbA + (yi - loA)) % 3));
          }
          ++jA;
        }
        if (tr) tr[3] = clock64();
        const int kb = k - lag;
        if (kb >= 0 && kb < nBin) {
          const int yi = loA + kb;  // B's input row == A's output row
          if (yi == loA && yi >= y0) {
            const int q = qbB + (yi - y0);
            mbar_wait(&temptyB[q % 3], (((uint32_t)(q / 3)) & 1u) ^ 1u);
          }
          if (yi + 1 < y1) {
            const int q = qbB + (yi + 1 - y0);
            mbar_wait(&temptyB[q % 3], (((uint32_t)(q / 3)) & 1u) ^ 1u);
          }
          if (tr) tr[4] = clock64();
          const int qo = qbA + kb;
          const int os = qo % D;
          mbar_wait(&ofull[os], ((uint32_t)(qo / D)) & 1u);
          tc_fence_after();
          if (tr) tr[5] = clock64();
          pair_issue_row<NCH, NP, false>(leader, tmem_base + (uint32_t)(3 * NP), aO_lo + (uint32_t)os * o_units, aO_hi, bB_lo, bB_hi, yi, y0, y1, qbB, loA);
          if (leader) {
            if (yi - 1 >= y0) umma_commit_addr(tfullB_s + 8u * (uint32_t)((qbB + (yi - 1 - y0)) % 3));
            if (yi == hiA - 1 && yi < y1) umma_commit_addr(tfullB_s + 8u * (uint32_t)((qbB + (yi - y0)) % 3));
          }
        }
        if (tr) tr[6] = clock64();
      }
      qbA += nBin;
      qbB += y1 - y0;
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------------------- conv A: accumulator -> activation -> ring row
    const int w = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;  // pixel of the 128-pixel segment
    int qb = 0, jb = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
      const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
      const bool inimg = s.cx * kPrStep + m < p.W;
      for (int r = loA + ((w - qb) & 1); r < hiA; r += 2) {
        const int q = qb + (r - loA);  // q % 2 == w
        const int slot = q % 3, os = q % D;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(slot * NP);
        long long* const te = (p.trace != nullptr && blockIdx.x == 1 && qd == 0 && lane == 0 && q < 128) ? p.trace + 8 * 128 + 8 * q : nullptr;
        if (te) te[0] = clock64();
        mbar_wait_parked(&tfullA[slot], ((uint32_t)(q / 3)) & 1u);
        tc_fence_after();
        if (te) te[1] = clock64();
        if (qd == 0 && lane == 0) {
          // every MMA up to input row min(r + 1, last) has completed: hand those stages back to the producer
          const int lo = r == loA ? iA0 : r + 1;
          const int hi = min(r + 1, iA1 - 1);
          for (int yi = lo; yi <= hi; ++yi) mbar_arrive(&empty[(jb + (yi - iA0)) % S]);
        }
        uint32_t acc[NCH][16];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) tmem_ld16(taddr + (uint32_t)(16 * ci), acc[ci]);
        tmem_ld_wait();
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) tmem_st16_zero(taddr + (uint32_t)(16 * ci));
        tmem_st_wait();
        // no tcgen05.fence here: wait::ld / wait::st have completed the accumulator accesses, and the fence would also
        // wait for this thread's global stores / loads still in flight (measured: 1000 instead of 400 cycles)
        __syncwarp();
        if (lane == 0) mbar_arrive(&temptyA[slot]);
        if (te) te[2] = clock64();
        mbar_wait_parked(&oempty[os], (((uint32_t)(q / D)) & 1u) ^ 1u);
        if (te) te[3] = clock64();
        uint8_t* const orow = oring + (size_t)os * kOrow + (size_t)(8 + m) * 16;
        if (!(p.dbg & 1))
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c0 = 16 * ci + 8 * half;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[i] = activate<true, ACTM>(ACTM, __uint_as_float(acc[ci][8 * half + i]) + biasA_sm[c0 + i], p.actA_param, 0.0f);
            uint4 o;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(v[0], v[1]);
            o.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[2], v[3]);
            o.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[4], v[5]);
            o.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[6], v[7]);
            o.w = *reinterpret_cast<uint32_t*>(&t);
            if (!inimg) o = make_uint4(0, 0, 0, 0);  // right of the image: B's zero padding
            *reinterpret_cast<uint4*>(orow + (size_t)(c0 >> 3) * kPrPlaneBytes) = o;
          }
        if (te) te[4] = clock64();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ofull[os]);
        if (te) te[5] = clock64();
      }
      qb += hiA - loA;
      jb += iA1 - iA0;
    }
  } else if (warp >= 12) {
    // ---------------------------------------------------------------- conv B: the usual fused tail -> global
    // Both warpgroups work on EVERY row, each on half of the channels (NCH granules of 8): a thread then holds 8*NCH
    // accumulators plus three rows' worth of residual chunks.  The residual of row y + 2 is requested right after row
    // y's accumulators have been read, two rows ahead of its use: loads still in flight when tcgen05.ld is waited for
    // would delay the hand-over of the accumulator slot (LDG and LDTM retire through the same scoreboard), and under
    // full HBM load a residual load takes > 2000 cycles.
    constexpr int G = NCH;  // granules per thread
    const int w = (warp - 12) >> 2;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    const int cstore = (p.epi.cout + 7) & ~7;
    constexpr bool kUsesRes = COMB == RSB_COMB_SPAB_GATE || COMB == RSB_COMB_MUL || COMB == RSB_COMB_AXPY;
    const size_t plane_stride = (size_t)p.H * p.W * 8;
    int qb = 0, qbA = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int y0 = s.y0, y1 = s.y1;
      const int loA = max(y0 - 1, 0), hiA = min(y1 + 1, p.H);
      const int n = s.n;
      const int x = s.cx * kPrStep + m;
      const int own_lo = s.cx == 0 ? 0 : s.cx * kPrStep + 1;
      const int own_hi = s.cx == p.cols - 1 ? p.W : s.cx * kPrStep + kPrStep + 1;
      const bool valid = x >= own_lo && x < own_hi && !(p.dbg & 2);
      uint4 pre[3][kUsesRes ? G : 1];
      auto fetch = [&](uint4* dst, int y) {
        if constexpr (kUsesRes) {
          if (valid && y < y1) {
            const T* rp = reinterpret_cast<const T*>(p.epi.res1) + planar_index(n, p.epi.res1_planes, p.epi.res1_plane0 + G * w, p.H, p.W, y, x);
#pragma unroll
            for (int g = 0; g < G; ++g)
              if ((G * w + g) * 8 < cstore) dst[g] = *reinterpret_cast<const uint4*>(rp + g * plane_stride);
          }
        }
      };
      fetch(pre[0], y0);
      fetch(pre[1], y0 + 1);
      for (int yb = y0; yb < y1; yb += 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int y = yb + k;
          if (y < y1) {
            const int q = qb + (y - y0);
            const int slot = q % 3;
            const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((3 + slot) * NP + 8 * G * w);
            long long* const te = (p.trace != nullptr && blockIdx.x == 1 && w == 0 && qd == 0 && lane == 0 && q < 128) ? p.trace + 8 * 256 + 8 * q : nullptr;
            if (te) te[0] = clock64();
            mbar_wait_parked(&tfullB[slot], ((uint32_t)(q / 3)) & 1u);
            tc_fence_after();
            if (te) te[1] = clock64();
            if (w == 0 && qd == 0 && lane == 0) {
              // B's MMAs up to ring row min(y + 1, last) have completed: those ring rows may be overwritten
              const int lo = y == y0 ? loA : y + 1;
              const int hi = min(y + 1, hiA - 1);
              for (int yy = lo; yy <= hi; ++yy) mbar_arrive(&oempty[(qbA + (yy - loA)) % D]);
            }
            uint32_t acc[G][8];
#pragma unroll
            for (int g = 0; g < G; ++g) tmem_ld8(taddr + (uint32_t)(8 * g), acc[g]);
            tmem_ld_wait();
            // (B's accumulators are not cleared: its MMAs overwrite on first touch)
            __syncwarp();
            if (lane == 0) mbar_arrive(&temptyB[slot]);
            if (te) te[2] = clock64();
            fetch(pre[(k + 2) % 3], y + 2);
            if (valid) {
              T* const drow = reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0, p.H, p.W, y, x);
#pragma unroll
              for (int g = 0; g < G; ++g)
                if ((G * w + g) * 8 < cstore)
                  epilogue8_planar<ACT, COMB>(p.epi, biasB_sm, biasB_sm, acc[g], (G * w + g) * 8, drow, plane_stride, kUsesRes ? &pre[k][g] : nullptr);
            }
            if (te) te[3] = clock64();
          }
        }
      }
      qb += y1 - y0;
      qbA += hiA - loA;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef void (*PairKernelFn)(const CUtensorMap, const CUtensorMap, const ConvPairParams);

struct PairVariant {
  int ks0, nch, actm, act, comb;
  PairKernelFn fn;
};

#define RSB_P(KS0, NCH, ACTM, ACT, COMB) {KS0, NCH, ACTM, ACT, COMB, conv_pair_kernel<KS0, NCH, ACTM, ACT, COMB>}
const PairVariant kPairVariants[] = {
    // SPAN / SPANPlus SPAB: c2_r (+ SiLU / Mish) -> c3_r + gate
    RSB_P(3, 3, RSB_ACT_SILU, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    RSB_P(3, 3, RSB_ACT_MISH, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    // plain pairs (tests, conv -> conv without a tail)
    RSB_P(3, 3, RSB_ACT_NONE, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_P(3, 3, RSB_ACT_SILU, RSB_ACT_SILU, RSB_COMB_NONE),
};
#undef RSB_P
constexpr int kNumPairVariants = sizeof(kPairVariants) / sizeof(kPairVariants[0]);

PairKernelFn pair_pick(int cin0, int np, int actA, int actB, int combB) {
  if (combB == RSB_COMB_SPAB_GATE) actB = RSB_ACT_NONE;  // the gate ignores `act`
  for (int i = 0; i < kNumPairVariants; ++i) {
    const PairVariant& v = kPairVariants[i];
    if (v.ks0 * 16 == cin0 && v.nch * 16 == np && v.actm == actA && v.act == actB && v.comb == combB) return v.fn;
  }
  return nullptr;
}

}  // namespace

uint32_t conv_pair_weight_bytes(int cin, int np) { return 3u * (uint32_t)(cin / 8) * (uint32_t)(5 * np) * 16u; }

size_t conv_pair_smem_bytes(int cin0, int np) {
  const uint32_t stage = (uint32_t)(cin0 / 8) * kPrPlaneBytes, orow = (uint32_t)(np / 8) * kPrPlaneBytes;
  return (size_t)pr_align(conv_pair_weight_bytes(cin0, np)) + pr_align(conv_pair_weight_bytes(np, np)) + (size_t)kPrStages * pr_align(stage) +
         (size_t)kPrDepth * orow + 2 * np * sizeof(float) + (2 * kPrStages + 2 * kPrDepth + 12 + 1) * 8 + 16;
}

int conv_pair_cols(int W) { return W <= 127 ? 1 : (W - 7 + kPrStep - 1) / kPrStep; }

bool conv_pair_supported(int cin0, int np, int actA, int actB, int combB) { return pair_pick(cin0, np, actA, actB, combB) != nullptr; }

cudaError_t conv_pair_configure(size_t max_smem) {
  for (int i = 0; i < kNumPairVariants; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kPairVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_pair(const CUtensorMap& src_map, const CUtensorMap& res_map, const ConvPairParams& p, int num_sms, cudaStream_t stream) {
  PairKernelFn fn = pair_pick(p.cin0, p.np, p.actA, p.epi.act, p.epi.combine);
  if (fn == nullptr) return cudaErrorInvalidValue;
  const size_t smem = conv_pair_smem_bytes(p.cin0, p.np);
  const int grid = p.units < num_sms ? p.units : num_sms;
  ConvPairParams q = p;
  static const char* lag_env = getenv("RSB_PAIR_LAG");
  static const char* dbg_env = getenv("RSB_PAIR_DBG");
  q.lag = lag_env != nullptr ? atoi(lag_env) : kPrLag;
  q.dbg = dbg_env != nullptr ? atoi(dbg_env) : 0;
  static const char* trace = getenv("RSB_PAIR_TRACE");
  if (trace != nullptr) {
    // bring-up: clock stamps of CTA 1 (MMA thread per step, A / B epilogue lane per row), dumped after every launch
    static long long* d_trace = nullptr;
    const size_t tbytes = 3 * 128 * 8 * sizeof(long long);
    if (d_trace == nullptr) cudaMalloc(&d_trace, tbytes);
    cudaMemsetAsync(d_trace, 0, tbytes, stream);
    q.trace = d_trace;
    cudaError_t e = launch_pdl(fn, dim3(grid), dim3(kPrThreads), smem, stream, src_map, res_map, q);
    static long long h[3 * 128 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, d_trace, tbytes, cudaMemcpyDeviceToHost);
    FILE* f = fopen(trace, "w");
    if (f != nullptr) {
      for (int r = 0; r < 3 * 128; ++r) {
        fprintf(f, "%d %d", r / 128, r % 128);
        for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[8 * r + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    return e;
  }
  return launch_pdl(fn, dim3(grid), dim3(kPrThreads), smem, stream, src_map, res_map, q);
}

}  // namespace rsb
