// conv_pair — TWO consecutive 3x3 stride-1 'same' convolutions (A, then B) fused into one row-streaming kernel.
//
// Why: an unfused 48->48 3x3 layer at 1080p moves 398 MB through HBM for 86 GFLOP — it sits exactly on the B200 ridge
// (profiles/r1b_rs_bound_switches.md: MMAs alone 63 us, loads + stores alone 76 us, together 79 us).  Here A's finished
// output rows go from TMEM through the epilogue warps straight into a shared-memory ring in the canonical K-major UMMA
// layout and are consumed there by B's MMAs: the intermediate map is never re-read from HBM (and not written either, unless a
// later layer needs it).  A plan's chain of 3x3 convs is cut into such pairs (plan.cu::find_pairs): SPAN's 20 convs become 10
// launches.
//
// Each conv is the row-streaming formulation of conv_rs.cu: one MMA multiplies a 128-pixel segment of ONE INPUT ROW by
// the weights of all three kernel rows, D[128 px x 3*NP] += X[y] * [W(kh=0) | W(kh=1) | W(kh=2)]; the three N blocks are the
// contributions of input row y to output rows y+1, y, y-1, whose accumulators sit side by side in a TMEM ring of NS = 5
// slots per conv in descending row order (a window that wraps is issued as two narrower MMAs; same weight layout as conv_rs).
// Differences to a single conv:
//   * strips advance by 120 pixels: B's outputs at the first and last pixel of a 128-pixel segment would need A's
//     outputs of the neighbouring strip, so strip cx owns the output pixels [120cx+1, 120cx+121) (image borders are
//     exact: the ring rows keep zero guard pixels = B's zero padding, and A's outputs right of the image are stored as
//     zeros).  Vertically a CTA streams a contiguous run of rows; A runs one row ahead/behind at the run's ends.
//   * the residual rows of a SPAB gate (conv B's, or conv A's) come through TMA into a small shared-memory ring: no
//     global load ever sits in an epilogue thread's scoreboard (round 1 read them with LDG: tcgen05.wait::ld then waited for
//     them too, and the accumulator hand-over the MMA thread polls for took ~1000 cycles).
//   * per output pixel the accumulation order is input rows y-1, y, y+1, each over (dx, 16-channel step) — identical
//     to conv_rs / conv_tc, and A's output is rounded to bf16 exactly as if it had been stored: the fused pair is
//     bit-identical to the two separate launches.
//
// Round 2 changes (tools/unit_times.py, tools/ubench/umma_dual.cu):
//   * EACH CONV HAS ITS OWN MMA-ISSUING WARP.  The tensor pipe retires one N = 144 MMA per 72 cycles whoever issues it, but
//     a thread that also polls three mbarriers and commits once per 9-MMA row only issues one per 137 cycles.
//   * five accumulator slots per conv instead of three.  With three, the MMAs of input row y + 1 need the slot that output
//     row y - 2 occupies, which completes with input row y - 1 and can only then be drained: MMA -> commit -> epilogue wake-up
//     -> tcgen05.ld -> clear -> hand-over -> MMA was one serial chain per row (2200 cycles per row pair measured, 1300 of MMA
//     work).  Five slots leave two rows of slack; the price is two split rows in five (+30 % tensor-pipe time per conv, which
//     the pipe has to spare: a pair is HBM-bound).
//
//   warp 0       TMA producer: A's input rows (ring of 4)
//   warp 1       tcgen05.mma issuer of conv A
//   warp 2       TMEM allocation, then TMA producer of the gate's residual rows (ring of 2)
//   warp 3       tcgen05.mma issuer of conv B
//   warps 4..11  two warpgroups draining A's accumulators (even / odd rows): bias + activation (or the SPAB gate with its
//                residual) -> bf16 -> st.shared into the ring (+ st.global when a later layer reads A's output)
//                -> fence.proxy.async -> mbarrier
//   warps 12..19 two warpgroups draining B's accumulators (even / odd rows): the usual fused epilogue (kernels.cuh) -> global
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kPrThreads = 640;
constexpr int kPrGroups = 18;  // 8-pixel groups per staged row: 1 left guard + 16 + 1 right guard
constexpr uint32_t kPrPlaneBytes = kPrGroups * 128u;
constexpr uint32_t kPrResPlaneBytes = 16u * 128u;  // residual rows: the 128 pixels of the segment, no guards
constexpr int kPrStages = 4;  // TMA ring of A's input rows
constexpr int kPrDepth = 4;   // ring of A's output rows
constexpr int kPrRes = 2;     // ring of residual rows (one per epilogue warpgroup)
constexpr int kPrSlots = 5;   // TMEM accumulator slots per conv
constexpr int kPrStep = 120;  // x distance between strips

__host__ __device__ inline uint32_t pr_align(uint32_t v) { return (v + 127u) & ~127u; }

// Bring-up builds (-DRSB_BRINGUP) record clock64 stamps of CTA 1's warps per row: [role][row][8]; roles: 0 MMA A, 1 MMA B,
// 2 A epilogue (warpgroup lane 0 of quarter 0), 3 B epilogue.  Read back with rsb_debug_pair_trace (tools/pair_trace.py).
#ifdef RSB_BRINGUP
constexpr int kTrRows = 96;
__device__ long long g_pair_trace[4 * kTrRows * 8];
#define PR_TRACE(role, row, k) \
  do { if (blockIdx.x == 1 && (row) < kTrRows) g_pair_trace[((role) * kTrRows + (row)) * 8 + (k)] = clock64(); } while (0)
#else
#define PR_TRACE(role, row, k) do { } while (0)
#endif

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

// MMAs of input row `yi` of a conv whose output rows [lo, hi) are live in this run; qbase = sequence number of output row lo.
// Output row with sequence number q lives in slot q % NS, slot s at TMEM column (NS - 1 - s) * NP of the conv's ring: rows
// yi+1, yi, yi-1 are three consecutive N blocks [kh = 0 | 1 | 2] unless the ring wraps or a row lies outside [lo, hi).
// Every accumulator slot is zero when it is handed over (the epilogue clears it after reading): all MMAs accumulate.
// Commits tfull of every output row whose last contribution this was (yi - 1 always, yi on the image's last row).
template <int KS, int NP, int NS>
__device__ __forceinline__ void ring_issue_row(bool leader, uint32_t tm, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo0, uint32_t b_hi, int yi, int lo,
                                               int hi, int H, int qbase, uint32_t tfull_s) {
  using namespace ptx;
  constexpr uint32_t kPU = kPrPlaneBytes >> 4;  // one 8-channel plane, in 16-byte units
  const int qn = qbase + (yi + 1 - lo);         // sequence number of output row yi + 1
  const int slot_n = qn % NS;
  if (yi - 1 >= lo && yi + 1 < hi && slot_n >= 2) {
    constexpr uint32_t kI3 = make_idesc_bf16(128, 3 * NP);
    const uint32_t d0 = tm + (uint32_t)(NS - 1 - slot_n) * NP;
    if (leader) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk)
          umma_bf16_lohi<true>(d0, a_lo + (uint32_t)dx + (uint32_t)kk * 2u * kPU, a_hi, b_lo0 + (uint32_t)((dx * 2 * KS + 2 * kk) * 3 * NP), b_hi, kI3);
      umma_commit_addr(tfull_s + 8u * (uint32_t)(slot_n - 2));  // output row yi - 1 is complete
    }
  } else {
    // run / image borders and ring wrap: N block g (kernel row kh = g) feeds output row r = yi + 1 - g
    constexpr uint32_t idesc0 = make_idesc_bf16(128, 0);
    bool v[3];
    uint32_t col[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int r = yi + 1 - g;
      v[g] = r >= lo && r < hi;
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
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          const uint32_t a = a_lo + (uint32_t)dx + (uint32_t)kk * 2u * kPU;
          const uint32_t b = b_lo0 + (uint32_t)((dx * 2 * KS + 2 * kk) * 3 * NP);
#pragma unroll
          for (int g = 0; g < 3; ++g)
            if (idg[g] != 0u) umma_bf16_lohi<true>(tm + col[g], a, a_hi, b + (uint32_t)(g * NP), b_hi, idg[g]);
        }
      if (v[2]) umma_commit_addr(tfull_s + 8u * (uint32_t)((qn - 2) % NS));
      if (v[1] && yi == H - 1) umma_commit_addr(tfull_s + 8u * (uint32_t)((qn - 1) % NS));
    }
  }
}

__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
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
  return o;
}

// KS0: conv A's input channels / 16; NCH: A's output = B's input = B's output channels / 16.
// Conv A's tail: ACTA (activation) or COMBA == RSB_COMB_SPAB_GATE (gate with residual); STOREA: A's rows are also written to
// global memory (a later layer reads them).  Conv B's tail: ACT, or COMB == RSB_COMB_SPAB_GATE.  At most one of the two
// convs has a gate; its residual rows arrive through res_map.
template <int KS0, int NCH, int ACTA, int COMBA, int STOREA, int ACT, int COMB>
__global__ void __launch_bounds__(kPrThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap res_map, const __grid_constant__ ConvPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  constexpr int NP = 16 * NCH;
  constexpr int S = kPrStages, D = kPrDepth, R = kPrRes, NS = kPrSlots;
  constexpr uint32_t kOrow = (uint32_t)(NP / 8) * kPrPlaneBytes;
  constexpr uint32_t kRrow = (uint32_t)(NP / 8) * kPrResPlaneBytes;
  constexpr bool kGateA = COMBA == RSB_COMB_SPAB_GATE;
  constexpr bool kGateB = COMB == RSB_COMB_SPAB_GATE;
  static_assert(!(kGateA && kGateB), "one residual ring: at most one gate per pair");
  static_assert(2 * NS * NP <= 512, "two accumulator rings must fit TMEM");
  constexpr bool kRes = kGateA || kGateB;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  const uint32_t wA_al = pr_align(p.wbytesA), wB_al = pr_align(p.wbytesB), st_al = pr_align(p.stage_bytes);
  uint8_t* const wA = smem;
  uint8_t* const wB = wA + wA_al;
  uint8_t* const stage0 = wB + wB_al;
  uint8_t* const oring = stage0 + (size_t)S * st_al;
  uint8_t* const rring = oring + (size_t)D * kOrow;
  float* const biasA_sm = reinterpret_cast<float*>(rring + (kRes ? (size_t)R * kRrow : 0));
  float* const biasB_sm = biasA_sm + NP;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(biasB_sm + NP);
  uint64_t* const full = bars;             // [S]   TMA -> MMA A (bytes landed AND the accumulator slot the row opens is free)
  uint64_t* const empty = full + S;        // [S]   A epilogue -> TMA
  uint64_t* const ofull = empty + S;       // [D]   A epilogue -> MMA B  (ring row written)
  uint64_t* const oempty = ofull + D;      // [D]   B epilogue -> A epilogue (ring row consumed)
  uint64_t* const rfull = oempty + D;      // [R]   TMA -> gate epilogue (residual row landed)
  uint64_t* const rempty = rfull + R;      // [R]   gate epilogue -> TMA
  uint64_t* const tfullA = rempty + R;     // [NS]  MMA A -> A epilogue
  uint64_t* const temptyA = tfullA + NS;   // [NS]  A epilogue -> TMA producer
  uint64_t* const tfullB = temptyA + NS;   // [NS]  MMA B -> B epilogue
  uint64_t* const temptyB = tfullB + NS;   // [NS]  B epilogue -> MMA B
  uint64_t* const wbar = temptyB + NS;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int s = 0; s < D; ++s) mbar_init(&ofull[s], 4), mbar_init(&oempty[s], 1);
    for (int s = 0; s < R; ++s) mbar_init(&rfull[s], 1), mbar_init(&rempty[s], 4);
    for (int a = 0; a < NS; ++a) {
      mbar_init(&tfullA[a], 1), mbar_init(&temptyA[a], 4);
      mbar_init(&tfullB[a], 1), mbar_init(&temptyB[a], 4);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
    // packed weights are not produced by the previous kernel: fetch them before the grid-dependency wait
    prefetch_tmap(&src_map);
    if (kRes) prefetch_tmap(&res_map);
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
    // all accumulator slots of both rings start out zero (afterwards the epilogues clear each slot they have read)
    const uint32_t tz = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < 2 * NS * NP; c += 16) tmem_st16_zero(tz + (uint32_t)c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  pdl_wait();  // everything below reads or writes activation buffers
  const int u0 = (int)((long long)p.units * blockIdx.x / gridDim.x);
  const int u1 = (int)((long long)p.units * (blockIdx.x + 1) / gridDim.x);
  const size_t plane_stride = (size_t)p.H * p.W * 8;

  if (warp == 0) {
    if (lane == 0) {
      int j = 0, qbA = 0, u = u0;
      Run s;
      while (next_run(p, u, u1, s)) {
        const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
        const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
        // A's input rows are pulled into L2 a few steps ahead: a TMA load then has to cover L2 latency only
        constexpr int kAhead = 6;
        for (int yp = iA0 + S; yp < min(iA1, iA0 + kAhead); ++yp) tma_prefetch_l2_5d(&src_map, 0, s.cx * (kPrStep / 8) - 1, yp, p.src_plane0, s.n);
        for (int yi = iA0; yi < iA1; ++yi, ++j) {
          if (yi + kAhead < iA1) tma_prefetch_l2_5d(&src_map, 0, s.cx * (kPrStep / 8) - 1, yi + kAhead, p.src_plane0, s.n);
          const int st = j % S;
          mbar_wait_parked(&empty[st], (((uint32_t)(j / S)) & 1u) ^ 1u);
          // A's output rows that receive their first contribution from this input row: yi + 1, and row 0 at yi == 0
          for (int r = (yi == 0 ? 0 : yi + 1); r <= yi + 1; ++r)
            if (r >= loA && r < hiA) {
              const int q = qbA + (r - loA);
              mbar_wait_parked(&temptyA[q % NS], (((uint32_t)(q / NS)) & 1u) ^ 1u);
            }
          mbar_expect_tx(&full[st], p.stage_bytes);
          tma_load_5d(stage0 + (size_t)st * st_al, &src_map, &full[st], 0, s.cx * (kPrStep / 8) - 1, yi, p.src_plane0, s.n);
        }
        qbA += hiA - loA;
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- residual rows of the gate -> shared memory (own warp: it is
    // paced by the gate's epilogue, which runs a whole pipeline depth behind A's input rows, and must not hold those back)
    if (kRes && lane == 0) {
      int kres = 0, u = u0;
      Run s;
      while (next_run(p, u, u1, s)) {
        const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
        // conv A's output rows (gate on A) or conv B's (gate on B)
        const int rlo = kGateA ? loA : s.y0, rhi = kGateA ? hiA : s.y1;
        constexpr int kAheadR = 4;  // L2 prefetch distance: the shared-memory load then covers L2 latency only
        for (int yp = rlo; yp < min(rhi, rlo + kAheadR); ++yp) tma_prefetch_l2_5d(&res_map, 0, s.cx * (kPrStep / 8), yp, p.res_plane0, s.n);
        for (int r = rlo; r < rhi; ++r, ++kres) {
          if (r + kAheadR < rhi) tma_prefetch_l2_5d(&res_map, 0, s.cx * (kPrStep / 8), r + kAheadR, p.res_plane0, s.n);
          const int rs = kres % R;
          mbar_wait_parked(&rempty[rs], (((uint32_t)(kres / R)) & 1u) ^ 1u);
          mbar_expect_tx(&rfull[rs], kRrow);
          tma_load_5d(rring + (size_t)rs * kRrow, &res_map, &rfull[rs], 0, s.cx * (kPrStep / 8), r, p.res_plane0, s.n);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- conv A: MMA issue
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint64_t dbA = make_smem_desc(smem_u32(wA), (uint32_t)(3 * NP) * 16u, 128u);
    const uint64_t daI = make_smem_desc(smem_u32(stage0) + 7u * 16u, kPrPlaneBytes, 128u);
    const uint32_t bA_lo = (uint32_t)dbA, bA_hi = (uint32_t)(dbA >> 32);
    const uint32_t aI_lo = (uint32_t)daI, aI_hi = (uint32_t)(daI >> 32);
    const uint32_t st_units = st_al >> 4;
    const uint32_t tfullA_s = smem_u32(tfullA);
    int jA = 0, qbA = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
      const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
      for (int yi = iA0; yi < iA1; ++yi, ++jA) {
        const int st = jA % S;
        if (leader) PR_TRACE(0, jA, 0);
        mbar_wait(&full[st], ((uint32_t)(jA / S)) & 1u);
        tc_fence_after();
        if (leader) PR_TRACE(0, jA, 1);
        ring_issue_row<KS0, NP, NS>(leader, tmem_base, aI_lo + (uint32_t)st * st_units, aI_hi, bA_lo, bA_hi, yi, loA, hiA, p.H, qbA, tfullA_s);
        if (leader) PR_TRACE(0, jA, 2);
      }
      qbA += hiA - loA;
    }
    __syncwarp();
  } else if (warp == 3) {
    // ---------------------------------------------------------------- conv B: MMA issue (input rows = A's output rows in the ring)
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint64_t dbB = make_smem_desc(smem_u32(wB), (uint32_t)(3 * NP) * 16u, 128u);
    const uint64_t daO = make_smem_desc(smem_u32(oring) + 7u * 16u, kPrPlaneBytes, 128u);
    const uint32_t bB_lo = (uint32_t)dbB, bB_hi = (uint32_t)(dbB >> 32);
    const uint32_t aO_lo = (uint32_t)daO, aO_hi = (uint32_t)(daO >> 32);
    const uint32_t o_units = kOrow >> 4;
    const uint32_t tfullB_s = smem_u32(tfullB);
    int qbA = 0, qbB = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int y0 = s.y0, y1 = s.y1;
      const int loA = max(y0 - 1, 0), hiA = min(y1 + 1, p.H);
      for (int yi = loA; yi < hiA; ++yi) {
        if (leader) PR_TRACE(1, qbA + (yi - loA), 0);
        // B's output rows that receive their first contribution from this input row: yi + 1, and row 0 at yi == 0
        for (int r = (yi == 0 ? 0 : yi + 1); r <= yi + 1; ++r)
          if (r >= y0 && r < y1) {
            const int q = qbB + (r - y0);
            mbar_wait(&temptyB[q % NS], (((uint32_t)(q / NS)) & 1u) ^ 1u);
          }
        const int qo = qbA + (yi - loA);
        const int os = qo % D;
        if (leader) PR_TRACE(1, qo, 1);
        mbar_wait(&ofull[os], ((uint32_t)(qo / D)) & 1u);
        tc_fence_after();
        if (leader) PR_TRACE(1, qo, 2);
        ring_issue_row<NCH, NP, NS>(leader, tmem_base + (uint32_t)(NS * NP), aO_lo + (uint32_t)os * o_units, aO_hi, bB_lo, bB_hi, yi, y0, y1, p.H, qbB,
                                    tfullB_s);
        if (leader) PR_TRACE(1, qo, 3);
      }
      qbA += hiA - loA;
      qbB += y1 - y0;
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------------------- conv A: accumulator -> tail -> ring row (+ global)
    const int w = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;  // pixel of the 128-pixel segment
    int qb = 0, jb = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int loA = max(s.y0 - 1, 0), hiA = min(s.y1 + 1, p.H);
      const int iA0 = max(loA - 1, 0), iA1 = min(hiA + 1, p.H);
      const int x = s.cx * kPrStep + m;
      const bool inimg = x < p.W;
      // rows of the run this CTA owns and pixels of the strip this CTA owns: only those go to global memory
      const int own_lo = s.cx == 0 ? 0 : s.cx * kPrStep + 1;
      const int own_hi = s.cx == p.cols - 1 ? p.W : s.cx * kPrStep + kPrStep + 1;
      const bool own_x = x >= own_lo && x < own_hi;
      for (int r = loA + ((w - qb) & 1); r < hiA; r += 2) {
        const int q = qb + (r - loA);  // q % 2 == w
        const int slot = q % NS, os = q % D;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((NS - 1 - slot) * NP);
        const bool tr = qd == 0 && lane == 0;
        if (tr) PR_TRACE(2, q, 0);
        mbar_wait_parked(&tfullA[slot], ((uint32_t)(q / NS)) & 1u);
        tc_fence_after();
        if (tr) PR_TRACE(2, q, 1);
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&temptyA[slot]);
        if (tr) PR_TRACE(2, q, 2);
        uint4 pre[kGateA ? 2 * NCH : 1];
        if constexpr (kGateA) {
          // residual row r: sequence number q, ring stage q % R (== this warpgroup's)
          const int rs = q % R;
          mbar_wait_parked(&rfull[rs], ((uint32_t)(q / R)) & 1u);
          const uint8_t* rrow = rring + (size_t)rs * kRrow + (size_t)m * 16;
#pragma unroll
          for (int k = 0; k < 2 * NCH; ++k) pre[k] = *reinterpret_cast<const uint4*>(rrow + (size_t)k * kPrResPlaneBytes);
          __syncwarp();
          if (lane == 0) mbar_arrive(&rempty[rs]);
        }
        if (tr) PR_TRACE(2, q, 3);
        mbar_wait_parked(&oempty[os], (((uint32_t)(q / D)) & 1u) ^ 1u);
        if (tr) PR_TRACE(2, q, 4);
        uint8_t* const orow = oring + (size_t)os * kOrow + (size_t)(8 + m) * 16;
        const bool to_global = STOREA != 0 && own_x && r >= s.y0 && r < s.y1;
        T* const grow = STOREA != 0 ? reinterpret_cast<T*>(p.dstA) + planar_index(s.n, p.dstA_planes, p.dstA_plane0, p.H, p.W, r, x) : nullptr;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c0 = 16 * ci + 8 * half;
            float v[8];
            const float4 b0 = reinterpret_cast<const float4*>(biasA_sm + c0)[0];
            const float4 b1 = reinterpret_cast<const float4*>(biasA_sm + c0)[1];
            v[0] = __uint_as_float(acc[ci][8 * half + 0]) + b0.x, v[1] = __uint_as_float(acc[ci][8 * half + 1]) + b0.y;
            v[2] = __uint_as_float(acc[ci][8 * half + 2]) + b0.z, v[3] = __uint_as_float(acc[ci][8 * half + 3]) + b0.w;
            v[4] = __uint_as_float(acc[ci][8 * half + 4]) + b1.x, v[5] = __uint_as_float(acc[ci][8 * half + 5]) + b1.y;
            v[6] = __uint_as_float(acc[ci][8 * half + 6]) + b1.z, v[7] = __uint_as_float(acc[ci][8 * half + 7]) + b1.w;
            if constexpr (kGateA) {
              float rr[8];
              unpack8<__nv_bfloat16>(pre[2 * ci + half], rr);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = spab_gate_fast(v[i], rr[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = activate<true, ACTA>(ACTA, v[i], p.actA_param, 0.0f);
            }
            uint4 o = pack8_bf16(v);
            if (!inimg) o = make_uint4(0, 0, 0, 0);  // right of the image: B's zero padding
            *reinterpret_cast<uint4*>(orow + (size_t)(c0 >> 3) * kPrPlaneBytes) = o;
            if constexpr (STOREA != 0) {
              if (to_global) *reinterpret_cast<uint4*>(grow + (size_t)(c0 >> 3) * plane_stride) = o;
            }
          }
        if (tr) PR_TRACE(2, q, 5);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ofull[os]);
        if (tr) PR_TRACE(2, q, 6);
      }
      qb += hiA - loA;
      jb += iA1 - iA0;
    }
  } else if (warp >= 12) {
    // ---------------------------------------------------------------- conv B: accumulator -> fused tail -> global
    const int w = (warp - 12) >> 2;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    const int cstore = (p.epi.cout + 7) & ~7;
    int qb = 0, qbA = 0, u = u0;
    Run s;
    while (next_run(p, u, u1, s)) {
      const int y0 = s.y0, y1 = s.y1;
      const int loA = max(y0 - 1, 0), hiA = min(y1 + 1, p.H);
      const int n = s.n;
      const int x = s.cx * kPrStep + m;
      const int own_lo = s.cx == 0 ? 0 : s.cx * kPrStep + 1;
      const int own_hi = s.cx == p.cols - 1 ? p.W : s.cx * kPrStep + kPrStep + 1;
      const bool valid = x >= own_lo && x < own_hi;
      for (int y = y0 + ((w - qb) & 1); y < y1; y += 2) {
        const int q = qb + (y - y0);  // q % 2 == w
        const int slot = q % NS;
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)((2 * NS - 1 - slot) * NP);
        const bool tr = qd == 0 && lane == 0;
        if (tr) PR_TRACE(3, q, 0);
        mbar_wait_parked(&tfullB[slot], ((uint32_t)(q / NS)) & 1u);
        tc_fence_after();
        if (tr) PR_TRACE(3, q, 1);
        if (qd == 0 && lane == 0) {
          // B's MMAs up to ring row min(y + 1, last) have completed: those ring rows may be overwritten
          const int lo = y == y0 ? loA : y + 1;
          const int hi = min(y + 1, hiA - 1);
          for (int yy = lo; yy <= hi; ++yy) mbar_arrive(&oempty[(qbA + (yy - loA)) % D]);
        }
        uint4 pre[kGateB ? 2 * NCH : 1];
        if constexpr (kGateB) {
          const int rs = q % R;
          mbar_wait_parked(&rfull[rs], ((uint32_t)(q / R)) & 1u);
          const uint8_t* rrow = rring + (size_t)rs * kRrow + (size_t)m * 16;
#pragma unroll
          for (int k = 0; k < 2 * NCH; ++k) pre[k] = *reinterpret_cast<const uint4*>(rrow + (size_t)k * kPrResPlaneBytes);
          __syncwarp();
          if (lane == 0) mbar_arrive(&rempty[rs]);
        }
        if (tr) PR_TRACE(3, q, 2);
        uint32_t r[2][16];
        tmem_ld16(taddr, r[0]);
        T* const drow = reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0, p.H, p.W, y, x);
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
            if (lane == 0) mbar_arrive(&temptyB[slot]);
            if (tr) PR_TRACE(3, q, 3);
          }
          if (valid) epilogue16_planar<ACT, COMB>(p.epi, biasB_sm, biasB_sm, r[ci & 1], c, cstore, drow, plane_stride, kGateB ? &pre[2 * ci] : nullptr);
        }
        if (tr) PR_TRACE(3, q, 4);
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
  int ks0, nch, acta, comba, storea, act, comb;
  PairKernelFn fn;
};

#define RSB_P(KS0, NCH, ACTA, COMBA, STOREA, ACT, COMB) \
  {KS0, NCH, ACTA, COMBA, STOREA, ACT, COMB, conv_pair_kernel<KS0, NCH, ACTA, COMBA, STOREA, ACT, COMB>}
#define RSB_P_FAMILY(ACTX)                                                                                               \
  /* stem -> c1_r (the stem's output also feeds conv_cat and the first gate) */                                          \
  RSB_P(1, 3, RSB_ACT_NONE, RSB_COMB_NONE, 1, ACTX, RSB_COMB_NONE),                                                      \
  /* c1_r -> c2_r, with and without c1_r's activated output kept (the last SPAB hands it to conv_cat) */                 \
  RSB_P(3, 3, ACTX, RSB_COMB_NONE, 0, ACTX, RSB_COMB_NONE),                                                              \
  RSB_P(3, 3, ACTX, RSB_COMB_NONE, 1, ACTX, RSB_COMB_NONE),                                                              \
  /* c2_r -> c3_r + gate */                                                                                              \
  RSB_P(3, 3, ACTX, RSB_COMB_NONE, 0, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),                                                 \
  /* c3_r + gate -> next block's c1_r (the block output is the next gate's residual: stored) */                          \
  RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE, 1, ACTX, RSB_COMB_NONE),                                                 \
  RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE, 0, ACTX, RSB_COMB_NONE)
const PairVariant kPairVariants[] = {
    RSB_P_FAMILY(RSB_ACT_SILU),  // SPAN
    RSB_P_FAMILY(RSB_ACT_MISH),  // SPANPlus
    // last block's c3_r + gate -> conv_2 (no activation)
    RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE, 1, RSB_ACT_NONE, RSB_COMB_NONE),
    // plain pairs (tests, conv -> conv without a tail)
    RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_NONE, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_P(3, 3, RSB_ACT_NONE, RSB_COMB_NONE, 1, RSB_ACT_NONE, RSB_COMB_NONE),
};
#undef RSB_P_FAMILY
#undef RSB_P
constexpr int kNumPairVariants = sizeof(kPairVariants) / sizeof(kPairVariants[0]);

PairKernelFn pair_pick(int cin0, int np, int actA, int combA, int storeA, int actB, int combB) {
  if (combA == RSB_COMB_SPAB_GATE) actA = RSB_ACT_NONE;  // the gate ignores `act`
  if (combB == RSB_COMB_SPAB_GATE) actB = RSB_ACT_NONE;
  for (int i = 0; i < kNumPairVariants; ++i) {
    const PairVariant& v = kPairVariants[i];
    if (v.ks0 * 16 == cin0 && v.nch * 16 == np && v.acta == actA && v.comba == combA && v.storea == (storeA ? 1 : 0) && v.act == actB && v.comb == combB)
      return v.fn;
  }
  return nullptr;
}

}  // namespace

size_t conv_pair_smem_bytes(int cin0, int np) {
  const uint32_t wA = 9u * (uint32_t)cin0 * (uint32_t)np * 2u, wB = 9u * (uint32_t)np * (uint32_t)np * 2u;
  const uint32_t stage = (uint32_t)(cin0 / 8) * kPrPlaneBytes, orow = (uint32_t)(np / 8) * kPrPlaneBytes, rrow = (uint32_t)(np / 8) * kPrResPlaneBytes;
  return (size_t)pr_align(wA) + pr_align(wB) + (size_t)kPrStages * pr_align(stage) + (size_t)kPrDepth * orow + (size_t)kPrRes * rrow +
         2 * np * sizeof(float) + (2 * kPrStages + 2 * kPrDepth + 2 * kPrRes + 4 * kPrSlots + 1) * 8 + 16;
}

int conv_pair_cols(int W) { return W <= 127 ? 1 : (W - 7 + kPrStep - 1) / kPrStep; }

bool conv_pair_supported(int cin0, int np, int actA, int combA, int storeA, int actB, int combB) {
  return pair_pick(cin0, np, actA, combA, storeA, actB, combB) != nullptr;
}

cudaError_t conv_pair_configure(size_t max_smem) {
  for (int i = 0; i < kNumPairVariants; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kPairVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_pair(const CUtensorMap& src_map, const CUtensorMap& res_map, const ConvPairParams& p, int num_sms, cudaStream_t stream) {
  PairKernelFn fn = pair_pick(p.cin0, p.np, p.actA, p.combA, p.dstA != nullptr, p.epi.act, p.epi.combine);
  if (fn == nullptr) return cudaErrorInvalidValue;
  // always more than half an SM's shared memory: one CTA per SM, so the 512-column TMEM allocation never contends
  size_t smem = conv_pair_smem_bytes(p.cin0, p.np);
  if (smem < 120 * 1024) smem = 120 * 1024;
  const int grid = p.units < num_sms ? p.units : num_sms;
  return launch_pdl(fn, dim3(grid), dim3(kPrThreads), smem, stream, src_map, res_map, p);
}

#ifdef RSB_BRINGUP
// bring-up builds only: copy the row trace of the last fused-pair launch (CTA 1) to the host; returns the number of long longs
extern "C" int rsb_debug_pair_trace(long long* host, int capacity) {
  const int n = 4 * kTrRows * 8;
  if (capacity < n) return -1;
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(host, g_pair_trace, n * sizeof(long long)) != cudaSuccess) return -2;
  return n;
}
#endif

}  // namespace rsb
