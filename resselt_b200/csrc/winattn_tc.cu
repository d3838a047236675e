// Shifted-window attention on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 plan, head-padded q / k / v.
//
// Reference math: DAT Spatial_Attention (/root/reference/resselt/archs/dat/arch.py:224-267, shift + mask :363-428,456-482, zero padding
// :443-449) and SwinIR WindowAttention (archs/swinir/arch.py:133-170,268-293).  The mma.sync kernel in dat_ops.cu spends ~11 issue
// slots per score (fragment shuffles, ldmatrix, per-chunk online-softmax corrections) at 16 resident warps per SM; here
//   * a tile is 128 queries of one (head, branch): half of a 256-token window (DAT 8x32 / 32x8), one 128-token window, or two
//     64-token windows (SwinIR 8x8; the two windows form a block-diagonal problem, the off-diagonal blocks of P are zeros);
//   * S = Q K^T is ONE tcgen05.mma pair (M = 128, N = keys of the tile, K = 32 = padded head_dim) into TMEM; q / k / v are staged
//     by 16-byte cp.async straight into the canonical no-swizzle core-matrix layout (planar-8 activations ARE that layout once
//     every head starts on a plane boundary: the caller pads heads to 32 channels with zero weight rows);
//   * the softmax runs with ONE QUERY ROW PER THREAD (tcgen05.ld 32x32b: lane = row): no shuffles, no online rescaling — pass 1
//     adds scale + position bias (+ shift mask) and keeps the row maximum, writing t back to TMEM; pass 2 turns t into
//     P = 2^(t - max) as bf16 pairs IN PLACE (P aliases the first half of the score columns);
//   * O = P V is issued with A = P read from TMEM and B = V through an MN-major descriptor (V staged exactly like K); dim 31 of every
//     V row is 1, so column 31 of O is the softmax denominator of the ROUNDED probabilities; O lands in the dead score columns.
// TMEM per CTA = the tile's key count (128 or 256 columns): 2-4 CTAs per SM overlap each other's staging / MMA waits.
#include <algorithm>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {
namespace {

using namespace ptx;

constexpr int kTQ = 128;  // queries per tile = TMEM lanes = threads per CTA
constexpr int kHP = 32;   // padded head width
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void sts_zero16(uint32_t saddr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(saddr), "r"(0u) : "memory");
}
// ld.shared with a compile-time byte offset folded into the instruction (an address computed in C++ costs one integer op per load)
template <int kOff>
__device__ __forceinline__ float lds_f32_off(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(saddr), "n"(kOff));
  return v;
}
template <int kOff>
__device__ __forceinline__ float2 lds_f32x2_off(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(saddr), "n"(kOff));
  return v;
}
// packed fp32 pairs (sm_100 FFMA2 / FADD2): two scores per issue slot
__device__ __forceinline__ void fma2(float& x0, float& x1, float s, float b0, float b1) {  // x = x * s + b
  asm("{\n\t"
      ".reg .b64 ra, rs, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rs, {%2, %2};\n\t"
      "mov.b64 rb, {%3, %4};\n\t"
      "fma.rn.f32x2 ra, ra, rs, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t"
      "}"
      : "+f"(x0), "+f"(x1)
      : "f"(s), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void add2(float& x0, float& x1, float c) {  // x = x + c
  asm("{\n\t"
      ".reg .b64 ra, rc;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rc, {%2, %2};\n\t"
      "add.rn.f32x2 ra, ra, rc;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t"
      "}"
      : "+f"(x0), "+f"(x1)
      : "f"(c));
}
// 32 lanes x 32 consecutive 32-bit columns -> TMEM
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
// 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A = 128 lanes x (K = 16 bf16 as 8 packed 32-bit columns)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
constexpr uint32_t kIdescBMajorMN = 1u << 16;  // B operand MN-major (V is [key][dim] with dim contiguous)

// Softmax of this thread's query row (one row per thread, tcgen05.ld 32x32b: lane = row), two passes over the score columns with
// the next chunk's tcgen05.ld in flight.  Pass 1: t = s * scale2 + bias (+ mask) goes back to TMEM in place, row maximum kept.
// Pass 2: P = 2^(t - max) as bf16 pairs.  Key j of the row's window sits in column j; keys are row-major over the window, kWs keys
// per window row.  The bias table is stored REVERSED (ascending with the key: R[(ky - qy + Hs - 1) * S + kx - qx + Ws - 1]) on an even
// row stride S, so the biases of keys (e, e + 1) are one 8-byte load at a compile-time offset from a per-thread base, and TWICE: copy A
// serves the threads whose pair index is even, copy B (60 bytes past a 128-byte boundary: 8-byte aligned for odd indices, and the
// other half of the bank window) the odd ones.  Scale + bias and the subtraction of the maximum run on packed fp32 pairs.
__host__ __device__ constexpr int tab_stride(int ws) { return ws == 8 ? 24 : (ws == 16 ? 48 : 64); }
__host__ __device__ constexpr int tab_copy_b_bytes(int hs, int ws) { return ((2 * hs - 1) * tab_stride(ws) * 4 + 127) / 128 * 128 + 60; }

template <int kWs, bool kMask, int kE>
struct ScoreStep {
  static __device__ __forceinline__ void run(uint32_t (&v)[32], uint32_t base, float scale2, const uint32_t (&labs)[8], int qlab, float& m0, float& m1) {
    constexpr int kOff = 4 * ((kE / kWs) * tab_stride(kWs) + (kE % kWs));
    const float2 b = lds_f32x2_off<kOff>(base);
    float t0 = __uint_as_float(v[kE]), t1 = __uint_as_float(v[kE + 1]);
    fma2(t0, t1, scale2, b.x, b.y);
    if (kMask) {
      if ((int)((labs[kE >> 2] >> (8 * (kE & 3))) & 0xFF) != qlab) t0 += -100.0f * kLog2e;
      if ((int)((labs[kE >> 2] >> (8 * ((kE + 1) & 3))) & 0xFF) != qlab) t1 += -100.0f * kLog2e;
    }
    if (kE & 2) m1 = fmaxf(fmaxf(m1, t0), t1); else m0 = fmaxf(fmaxf(m0, t0), t1);
    v[kE] = __float_as_uint(t0), v[kE + 1] = __float_as_uint(t1);
    ScoreStep<kWs, kMask, kE + 2>::run(v, base, scale2, labs, qlab, m0, m1);
  }
};
template <int kWs, bool kMask>
struct ScoreStep<kWs, kMask, 32> {
  static __device__ __forceinline__ void run(uint32_t (&)[32], uint32_t, float, const uint32_t (&)[8], int, float&, float&) {}
};

template <int kWs, bool kMask>
__device__ __forceinline__ float score_chunk(uint32_t (&v)[32], int i, uint32_t qaddr, float scale2, const uint8_t* klab, int qlab) {
  const uint32_t base = qaddr + 4u * (uint32_t)(i * (32 / kWs) * tab_stride(kWs));
  uint32_t labs[8] = {};
  if (kMask) {
#pragma unroll
    for (int e = 0; e < 8; ++e) labs[e] = reinterpret_cast<const uint32_t*>(klab + 32 * i)[e];
  }
  float m0 = -INFINITY, m1 = -INFINITY;
  ScoreStep<kWs, kMask, 0>::run(v, base, scale2, labs, qlab, m0, m1);
  return fmaxf(m0, m1);
}

// Pass 1 over ncols (a multiple of 64) columns; the last load issued is chunk 0 again, for pass 2 (va).
template <int kWs, bool kMask>
__device__ __forceinline__ float softmax_max(uint32_t taddr, int ncols, uint32_t qaddr, float scale2, const uint8_t* klab, int qlab, uint32_t (&va)[32]) {
  const int nch = ncols / 32;
  uint32_t vb[32];
  float mx = -INFINITY;
  tmem_ld32(taddr, va);
  for (int i = 0; i < nch; i += 2) {
    tmem_ld_wait();
    tmem_ld32(taddr + 32 * (i + 1), vb);
    mx = fmaxf(mx, score_chunk<kWs, kMask>(va, i, qaddr, scale2, klab, qlab));
    tmem_st32(taddr + 32 * i, va);
    tmem_ld_wait();
    if (i + 2 < nch) tmem_ld32(taddr + 32 * (i + 2), va);
    mx = fmaxf(mx, score_chunk<kWs, kMask>(vb, i + 1, qaddr, scale2, klab, qlab));
    tmem_st32(taddr + 32 * (i + 1), vb);
  }
  tmem_st_wait();
  tmem_ld32(taddr, va);
  return mx;
}

// Pass 2: P = 2^(t - max) as bf16 pairs to paddr (may alias the first half of the score columns: chunk i reads columns
// [32 i, 32 i + 32) before it writes [16 i, 16 i + 16)); chunk 0 is already in flight into va.
__device__ __forceinline__ uint32_t exp_pair(uint32_t a, uint32_t b, float nmx) {
  float t0 = __uint_as_float(a), t1 = __uint_as_float(b);
  add2(t0, t1, nmx);
  return pack_bf16x2(ex2_approx(t0), ex2_approx(t1));
}
__device__ __forceinline__ void softmax_exp(uint32_t taddr, uint32_t paddr, int ncols, float mx, uint32_t (&va)[32]) {
  const int nch = ncols / 32;
  const float nmx = -mx;
  uint32_t vb[32];
  for (int i = 0; i < nch; i += 2) {
    uint32_t pk[16];
    tmem_ld_wait();
    tmem_ld32(taddr + 32 * (i + 1), vb);
#pragma unroll
    for (int e = 0; e < 16; ++e) pk[e] = exp_pair(va[2 * e], va[2 * e + 1], nmx);
    tmem_ld_wait();
    tmem_st16(paddr + 16 * i, pk);
    if (i + 2 < nch) tmem_ld32(taddr + 32 * (i + 2), va);
#pragma unroll
    for (int e = 0; e < 16; ++e) pk[e] = exp_pair(vb[2 * e], vb[2 * e + 1], nmx);
    tmem_st16(paddr + 16 * (i + 1), pk);
  }
}

// Launch-invariant geometry (window sides and token counts are powers of two: divisions by them are shifts)
struct Geometry {
  int N, lN, TK, tiles, hpb, per;  // tokens per window (and log2), keys per tile, tiles per image, heads per branch, tiles * n
};

struct Item {
  int br, h, n, tile;
};
__device__ __forceinline__ Item decode_item(const Geometry& g, int item) {
  Item t;
  const int bh = item / g.per, rem = item - bh * g.per;
  t.br = bh / g.hpb, t.h = bh - t.br * g.hpb, t.n = rem / g.tiles, t.tile = rem - t.n * g.tiles;
  return t;
}

// floats of the strided bias table: the larger of the two branches' (2 Hs - 1) rows x stride(Ws)
__host__ __device__ inline int tab_floats(int sh, int sw) {
  const int a = tab_copy_b_bytes(sh, sw) + (2 * sh - 1) * tab_stride(sw) * 4, b = tab_copy_b_bytes(sw, sh) + (2 * sw - 1) * tab_stride(sh) * 4;
  return ((a > b ? a : b) / 4 + 4) & ~3;
}

// per-buffer shared memory: Q [4 planes][128][8] | K [4][TK][8] | V [4][TK][8] | key labels [TK] | query info [128] (po, lab | ty << 8 | tx << 16)
__host__ __device__ inline uint32_t buffer_bytes(int TK) { return 4 * kTQ * 16 + 8 * (uint32_t)TK * 16 + (uint32_t)TK + kTQ * 8; }

// Stage one tile with 16-byte cp.async: thread t owns key token t (kThreads == TK); the thread whose token is query row r also
// stages Q row r and publishes the row's pixel offset / label / window coordinates.  Returns "a key label differs from its window's".
__device__ __forceinline__ int stage_tile(const WinAttnParams& p, const Geometry& g, const Item& it, uint8_t* buf) {
  using T = __nv_bfloat16;
  const int Hs = it.br == 0 ? p.split_h : p.split_w, Ws = it.br == 0 ? p.split_w : p.split_h;
  const int lW = 31 - __clz(Ws);
  const int sh = p.shifted ? Hs >> 1 : 0, sw = p.shifted ? Ws >> 1 : 0;
  const int nWx = p.Wp >> lW, nWin = nWx * (p.Hp / Hs);
  const int N = g.N, TK = g.TK;
  // first window of the tile, which half of it the queries are (256-token windows)
  const int win0 = N > kTQ ? it.tile >> 1 : it.tile << (7 - g.lN);
  const int qoff = N > kTQ ? (it.tile & 1) * kTQ : 0;
  const size_t hw = (size_t)p.H * p.W;
  const int ch0 = p.src_ch_off + (it.br * g.hpb + it.h) * kHP;
  const size_t plane_stride = hw * 8;
  const T* base = reinterpret_cast<const T*>(p.src) + ((size_t)it.n * p.src_planes + (ch0 >> 3)) * plane_stride;
  const size_t part = (size_t)(p.qkv_stride >> 3) * plane_stride;
  const uint32_t q_s = smem_u32(buf), k_s = q_s + 4 * kTQ * 16, v_s = k_s + 4 * (uint32_t)TK * 16;
  uint8_t* klab = buf + 4 * kTQ * 16 + 8 * TK * 16;
  int2* qinfo = reinterpret_cast<int2*>(klab + TK);
  int differs = 0;
  for (int kt = threadIdx.x; kt < TK; kt += blockDim.x) {
    const int win = win0 + (kt >> g.lN), t = kt & (N - 1);
    const int ty = t >> lW, tx = t & (Ws - 1);
    int po = -1, lab = 0;
    if (win < nWin) {
      const int wy = win / nWx, wx = win - wy * nWx;
      const int y0 = wy * Hs, x0 = wx * Ws;  // window origin in the rolled, padded image
      const int yr = y0 + ty, xr = x0 + tx;
      int yo = yr + sh, xo = xr + sw;        // where the token lives in the un-rolled image
      if (yo >= p.Hp) yo -= p.Hp;
      if (xo >= p.Wp) xo -= p.Wp;
      if (p.shifted) {
        const int by = p.Hp - Hs, bx = p.Wp - Ws;
        lab = 3 * (yr < by ? 0 : (yr < p.Hp - sh ? 1 : 2)) + (xr < bx ? 0 : (xr < p.Wp - sw ? 1 : 2));
        const int lab0 = 3 * (y0 < by ? 0 : 1) + (x0 < bx ? 0 : 1);  // the window's first token
        differs |= lab != lab0;
      }
      if (yo < p.H && xo < p.W) po = yo * p.W + xo;
    }
    const T* qg = base + (size_t)(po < 0 ? 0 : po) * 8;
    klab[kt] = (uint8_t)lab;
    const int r = kt - qoff;
    const bool is_q = r >= 0 && r < kTQ;
    if (is_q) qinfo[r] = make_int2(po, lab | (ty << 8) | (tx << 16));
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
      const uint32_t ko = (uint32_t)(pl * TK + kt) * 16u, qo = (uint32_t)(pl * kTQ + r) * 16u;
      if (po >= 0) {
        const T* gq = qg + pl * plane_stride;
        cp_async16(k_s + ko, gq + part);
        cp_async16(v_s + ko, gq + 2 * part);
        if (is_q) cp_async16(q_s + qo, gq);
      } else {
        sts_zero16(k_s + ko);
        sts_zero16(v_s + ko);
        if (is_q) sts_zero16(q_s + qo);
      }
    }
  }
  cp_async_commit();
  return differs;
}

// Persistent CTAs (512 / TK per SM: TMEM is the limit) of kThreads == TK threads, each walking items b, b + grid, ...: the q / k / v
// of the next item are in flight (cp.async into the other shared-memory buffer) while this item's softmax runs; TMEM / barriers are
// set up once, the bias table again when the head changes.  256-token windows run with two warpgroups on the same 128 rows: warpgroup
// c takes key columns [128 c, 128 c + 128) (row maxima exchanged through shared memory) — twice the warps to hide the softmax's
// dependent-issue latencies behind.
template <int kThreads>
__global__ void __launch_bounds__(kThreads, 512 / kThreads) winattn_tc_kernel(const __grid_constant__ WinAttnParams p, int total_items) {
  extern __shared__ __align__(128) uint8_t smraw[];
  using T = __nv_bfloat16;
  constexpr int TK = kThreads;
  Geometry g;
  g.N = p.split_h * p.split_w, g.lN = 31 - __clz(g.N), g.TK = TK;
  {
    const int nWin = (p.Wp / p.split_w) * (p.Hp / p.split_h);
    g.tiles = g.N > kTQ ? nWin * 2 : (nWin * g.N + kTQ - 1) / kTQ;
  }
  g.hpb = p.heads / 2, g.per = g.tiles * p.n;
  const int N = g.N;
  const int d = p.head_dim;
  const int tab_n = (2 * p.split_h - 1) * (2 * p.split_w - 1);
  const uint32_t buf_bytes = buffer_bytes(TK);
  uint8_t* bufs = smraw;
  float* tab = reinterpret_cast<float*>(smraw + 2 * buf_bytes);  // bias of the current head x log2 e
  float* rowmax = tab + tab_floats(p.split_h, p.split_w);        // [2][128] (two-warpgroup variant)
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowmax + 2 * kTQ);  // S ready, O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  const int row = threadIdx.x & (kTQ - 1), half = threadIdx.x >> 7;  // query row, column half (0 for 128-thread CTAs)

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, (uint32_t)TK);
    tmem_relinquish();
  }
  int item = blockIdx.x;
  Item cur = decode_item(g, item), nxt = cur;
  int differs = stage_tile(p, g, cur, bufs), differs_next = 0;
  int tab_bh = -1;
  const float scale2 = p.scale * kLog2e;
  // columns this thread scores.  64-token windows share a tile in pairs (rows 0-63 | 64-127 <-> columns 0-63 | 64-127);
  // two-warpgroup CTAs (256 keys) split the columns in halves, P of half c goes IN PLACE to the start of that half
  const int blk = N < kTQ ? row >> g.lN : 0;
  const int ncols = kThreads == 256 ? 128 : (N < kTQ ? N : TK);
  const int col0 = kThreads == 256 ? 128 * half : blk * N;
  const int pcol0 = kThreads == 256 ? 128 * half : blk * N / 2;
  const uint32_t o_col = kThreads == 256 ? 64u : (uint32_t)TK / 2;  // O in dead score columns

  for (int it = 0; item < total_items; item += gridDim.x, ++it) {
    uint8_t* buf = bufs + (size_t)(it & 1) * buf_bytes;
    const bool more = item + (int)gridDim.x < total_items;
    // the other buffer's last readers (the MMAs of the previous item) have completed: every thread waited on bars[1]
    if (more) {
      nxt = decode_item(g, item + gridDim.x);
      differs_next = stage_tile(p, g, nxt, bufs + (size_t)((it + 1) & 1) * buf_bytes);
    }
    const int Hs = cur.br == 0 ? p.split_h : p.split_w, Ws = cur.br == 0 ? p.split_w : p.split_h;
    const int tab_w = 2 * Ws - 1, tab_s = tab_stride(Ws);
    if (cur.br * g.hpb + cur.h != tab_bh) {  // (no thread is still in the previous item's softmax: that ended before its PV MMA)
      tab_bh = cur.br * g.hpb + cur.h;
      const float* table = cur.br == 0 ? p.table0 : p.table1;
      float* tab_b = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tab) + tab_copy_b_bytes(Hs, Ws));
      for (int i = threadIdx.x; i < tab_n; i += kThreads) {  // reversed, on the even row stride, both copies
        const int dy = i / tab_w, dx = i - dy * tab_w;
        const int r = (2 * Hs - 2 - dy) * tab_s + (2 * Ws - 2 - dx);
        const float v = table[(size_t)i * g.hpb + cur.h] * kLog2e;  // softmax in base 2
        tab[r] = v, tab_b[r] = v;
      }
    }
    if (more) cp_async_wait_group1(); else cp_async_wait_all();
    // dim 31 of every V row = 1: column 31 of O = sum of the rounded probabilities (head_dim < 32 always here)
    *reinterpret_cast<T*>(buf + 4 * kTQ * 16 + (size_t)(4 * TK + 3 * TK + (int)threadIdx.x) * 16 + 14) = __float2bfloat16_rn(1.0f);
    fence_proxy_async_smem();  // generic-proxy writes (cp.async, st.shared) -> visible to the tensor core's async-proxy reads
    tc_fence_before();         // (and the previous item's tcgen05.ld of O precede this item's MMA)
    const bool mixed = __syncthreads_or(differs) != 0;  // the shift mask only exists in windows that straddle the roll seam
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t q_s = smem_u32(buf), k_s = q_s + 4 * kTQ * 16, v_s = k_s + 4 * (uint32_t)TK * 16;
    const uint8_t* klab = buf + 4 * kTQ * 16 + 8 * TK * 16;
    const int2 qi = reinterpret_cast<const int2*>(klab + TK)[row];
    const int qlab = qi.y & 0xFF, qty = (qi.y >> 8) & 0xFF, qtx = qi.y >> 16;

    // ---- S = Q K^T: two K = 16 steps of one M = 128, N = TK MMA
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_bf16(kTQ, TK);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t da = make_smem_desc(q_s + (uint32_t)(ks * 2 * kTQ * 16), kTQ * 16u, 128u);
        const uint64_t db = make_smem_desc(k_s + (uint32_t)(ks * 2 * TK * 16), (uint32_t)TK * 16u, 128u);
        umma_bf16(tmem_base, da, db, idesc, ks);
      }
      umma_commit(&bars[0]);
    }
    // bias address of this row's first column (window row col0 / Ws of the keys): tab[qpos - ky * tab_w - kx]
    // index of the pair (key 0, key 1) of this row's first column: even for odd qtx (copy A), odd for even qtx (copy B)
    const int ky0 = kThreads == 256 ? (128 * half) >> (31 - __clz(Ws)) : 0;
    const int j0 = (Hs - 1 - qty + ky0) * tab_s + (Ws - 1 - qtx);
    const uint32_t qaddr = smem_u32(tab) + ((qtx & 1) ? 0u : (uint32_t)tab_copy_b_bytes(Hs, Ws)) + 4u * (uint32_t)j0;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t taddr = lane_base + (uint32_t)col0, paddr = lane_base + (uint32_t)pcol0;
    const uint8_t* klab_row = klab + col0;
    mbar_wait(&bars[0], it & 1);
    tc_fence_after();

    uint32_t va[32];
    float mx;
    if (Ws == 32) mx = mixed ? softmax_max<32, true>(taddr, ncols, qaddr, scale2, klab_row, qlab, va) : softmax_max<32, false>(taddr, ncols, qaddr, scale2, klab_row, 0, va);
    else if (Ws == 16) mx = mixed ? softmax_max<16, true>(taddr, ncols, qaddr, scale2, klab_row, qlab, va) : softmax_max<16, false>(taddr, ncols, qaddr, scale2, klab_row, 0, va);
    else mx = mixed ? softmax_max<8, true>(taddr, ncols, qaddr, scale2, klab_row, qlab, va) : softmax_max<8, false>(taddr, ncols, qaddr, scale2, klab_row, 0, va);
    if (kThreads == 256) {  // the row's other half
      rowmax[half * kTQ + row] = mx;
      __syncthreads();
      mx = fmaxf(mx, rowmax[(1 - half) * kTQ + row]);
    }
    softmax_exp(taddr, paddr, ncols, mx, va);
    if (N < kTQ) {  // the other window's block of P (64-token windows): zeros
      const uint32_t other = lane_base + (uint32_t)((1 - blk) * N / 2);
      for (int c = 0; c < N / 2; c += 16) tmem_st16_zero(other + c);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();

    // ---- O = P V: A = P from TMEM (8 packed columns per K = 16 step), B = V [key][dim] MN-major
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(kTQ, kHP) | kIdescBMajorMN;
      for (int kk = 0; kk < TK / 16; ++kk) {
        // MN-major, no swizzle: 8 keys x 16 bytes (8 dims) = one core matrix; LBO = next 8 keys (128 B), SBO = next 8 dims (plane)
        const uint64_t db = make_smem_desc(v_s + (uint32_t)kk * 256u, 128u, (uint32_t)TK * 16u);
        // P of keys [128 c, 128 c + 128) sits at columns [128 c, 128 c + 64) in the two-warpgroup layout, else contiguous
        const uint32_t a_col = kThreads == 256 ? (uint32_t)(kk >> 3) * 128u + (uint32_t)(kk & 7) * 8u : (uint32_t)kk * 8u;
        umma_bf16_ts(tmem_base + o_col, tmem_base + a_col, db, idesc, kk);
      }
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], it & 1);
    tc_fence_after();
    {
      uint32_t o[32];
      tmem_ld32(lane_base + o_col, o);
      tmem_ld_wait();
      if (qi.x >= 0) {
        const size_t hw = (size_t)p.H * p.W;
        const float inv = 1.0f / __uint_as_float(o[31]);
        const int co = p.dst_ch_off + (cur.br * g.hpb + cur.h) * kHP;
        T* out = reinterpret_cast<T*>(p.dst) + (((size_t)cur.n * p.dst_planes + (co >> 3)) * hw + qi.x) * 8;
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
          if (kThreads == 256 && (pl >> 1) != half) continue;  // two warpgroups: two planes each
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = pl * 8 + 2 * e;
            const float a = c < d ? __uint_as_float(o[c]) * inv : 0.0f, b = c + 1 < d ? __uint_as_float(o[c + 1]) * inv : 0.0f;
            w[e] = pack_bf16x2(a, b);
          }
          *reinterpret_cast<uint4*>(out + (size_t)pl * hw * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    cur = nxt, differs = differs_next;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, (uint32_t)TK);
}

size_t winattn_tc_smem_bytes(int split_h, int split_w) {
  const int N = split_h * split_w, TK = std::max(N, kTQ);
  return (size_t)2 * buffer_bytes(TK) + (size_t)tab_floats(split_h, split_w) * 4 + 2 * kTQ * 4 + 2 * 8 + 16;
}

}  // namespace

bool winattn_tc_supported(int heads, int head_dim, int split_h, int split_w) {
  const int N = split_h * split_w;
  return heads % 2 == 0 && head_dim < kHP && (N == 64 || N == 128 || N == 256) && split_h % 8 == 0 && split_w % 8 == 0 && split_h <= 32 && split_w <= 32;
}

cudaError_t winattn_tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(winattn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(winattn_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
}

cudaError_t launch_winattn_tc(const WinAttnParams& p, int num_sms, cudaStream_t s) {
  const int N = p.split_h * p.split_w;
  const int nWin = (p.Hp / p.split_h) * (p.Wp / p.split_w);
  const int tiles = N > kTQ ? nWin * (N / kTQ) : (nWin * N + kTQ - 1) / kTQ;
  const int total = tiles * p.n * p.heads;  // (branch, head) major
  const int per_sm = 512 / std::max(N, kTQ);  // resident CTAs per SM: TMEM columns
  const int grid = std::min(total, num_sms * per_sm);
  // (a 128-thread / 128-column variant for 256-token windows — the keys in two blocks merged in registers like split-K flash
  // attention, four CTAs per SM — is correct and measured 285 us per launch against 266 us for the two-warpgroup form below: the same
  // sixteen warps per SM, two more MMA round trips per tile)
  if (N > kTQ)
    winattn_tc_kernel<256><<<grid, 256, winattn_tc_smem_bytes(p.split_h, p.split_w), s>>>(p, total);
  else
    winattn_tc_kernel<128><<<grid, 128, winattn_tc_smem_bytes(p.split_h, p.split_w), s>>>(p, total);
  return cudaGetLastError();
}

}  // namespace rsb
