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
__device__ __forceinline__ void sts_zero16(uint32_t saddr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(saddr), "r"(0u) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
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

// Pass 1 over this thread's query row: t = s * scale2 + bias (+ mask), row maximum; t goes back to TMEM.
// Key j of the row's window sits in column j; keys are row-major over the window, kWs keys per window row, so the bias
// addresses of a 32-key chunk are compile-time offsets from one base: tab[qpos - ky * (2 kWs - 1) - kx].
template <int kWs, bool kMask>
__device__ __forceinline__ float softmax_pass1(uint32_t taddr, int ncols, uint32_t qaddr, float scale2, const uint8_t* klab, int qlab) {
  constexpr int kTabW = 2 * kWs - 1;
  float mx = -INFINITY;
  for (int i = 0; i < ncols / 32; ++i) {
    uint32_t v[32];
    tmem_ld32(taddr + 32 * i, v);
    tmem_ld_wait();
    const uint32_t base = qaddr - 4u * (uint32_t)(i * (32 / kWs) * kTabW);
    uint32_t labs[8];
    if (kMask) {
#pragma unroll
      for (int e = 0; e < 8; ++e) labs[e] = reinterpret_cast<const uint32_t*>(klab + 32 * i)[e];
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int off = (e / kWs) * kTabW + (e % kWs);
      float t = fmaf(__uint_as_float(v[e]), scale2, lds_f32(base - 4u * (uint32_t)off));
      if (kMask) {
        const int lab = (labs[e >> 2] >> (8 * (e & 3))) & 0xFF;
        if (lab != qlab) t += -100.0f * kLog2e;
      }
      mx = fmaxf(mx, t);
      v[e] = __float_as_uint(t);
    }
    tmem_st32(taddr + 32 * i, v);
  }
  tmem_st_wait();
  return mx;
}

// Pass 2: P = 2^(t - max) as bf16 pairs; pcol may alias the first half of the t columns (chunk i reads columns [32 i, 32 i + 32)
// before it writes [16 i, 16 i + 16)).
__device__ __forceinline__ void softmax_pass2(uint32_t taddr, uint32_t paddr, int ncols, float mx) {
  for (int i = 0; i < ncols / 32; ++i) {
    uint32_t v[32];
    tmem_ld32(taddr + 32 * i, v);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 16; ++e)
      pk[e] = pack_bf16x2(ex2_approx(__uint_as_float(v[2 * e]) - mx), ex2_approx(__uint_as_float(v[2 * e + 1]) - mx));
    tmem_st16(paddr + 16 * i, pk);
  }
}

struct TokenPos {
  int po;   // pixel index y * W + x in the un-rolled image, -1: padding (q = k = v = 0) or a window past the last one
  int lab;  // shift-mask region label
  int ty, tx;
};

__device__ __forceinline__ TokenPos token_pos(const WinAttnParams& p, int Hs, int Ws, int nWx, int nWin, int win, int t, int sh, int sw) {
  TokenPos r;
  r.ty = t / Ws, r.tx = t - r.ty * Ws;
  r.lab = 0, r.po = -1;
  if (win >= nWin) return r;
  const int wy = win / nWx, wx = win - wy * nWx;
  const int yr = wy * Hs + r.ty, xr = wx * Ws + r.tx;  // coordinates in the rolled, padded image
  int yo = yr + sh, xo = xr + sw;                      // where the token lives in the un-rolled image
  if (yo >= p.Hp) yo -= p.Hp;
  if (xo >= p.Wp) xo -= p.Wp;
  if (p.shifted) {
    const int ry = yr < p.Hp - Hs ? 0 : (yr < p.Hp - sh ? 1 : 2);
    const int rx = xr < p.Wp - Ws ? 0 : (xr < p.Wp - sw ? 1 : 2);
    r.lab = 3 * ry + rx;
  }
  if (yo < p.H && xo < p.W) r.po = yo * p.W + xo;
  return r;
}

__global__ void __launch_bounds__(kTQ) winattn_tc_kernel(const __grid_constant__ WinAttnParams p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  using T = __nv_bfloat16;
  const int br = blockIdx.z, h = blockIdx.y;
  const int Hs = br == 0 ? p.split_h : p.split_w, Ws = br == 0 ? p.split_w : p.split_h;
  const int N = Hs * Ws;
  const int TK = N > kTQ ? N : kTQ;  // keys of the tile (128 or 256)
  const int sh = p.shifted ? Hs / 2 : 0, sw = p.shifted ? Ws / 2 : 0;
  const int nWx = p.Wp / Ws, nWy = p.Hp / Hs, nWin = nWx * nWy;
  const int tiles = N > kTQ ? nWin * (N / kTQ) : (nWin * N + kTQ - 1) / kTQ;
  const int tile = blockIdx.x % tiles, n = blockIdx.x / tiles;
  // first window of the tile, which half of it the queries are (256-token windows)
  const int win0 = N > kTQ ? tile / (N / kTQ) : tile * (kTQ / N);
  const int qoff = N > kTQ ? (tile % (N / kTQ)) * kTQ : 0;
  const int tab_w = 2 * Ws - 1, tab_n = (2 * Hs - 1) * tab_w;
  const int d = p.head_dim, hpb = p.heads / 2;

  uint8_t* Qs = smraw;                       // [4 planes][128 tokens][8 dims] bf16
  uint8_t* Ks = Qs + 4 * kTQ * 16;           // [4][TK][8]
  uint8_t* Vs = Ks + 4 * TK * 16;            // [4][TK][8]
  float* tab = reinterpret_cast<float*>(Vs + 4 * TK * 16);  // bias of this head x log2 e
  uint8_t* klab = reinterpret_cast<uint8_t*>(tab + ((tab_n + 3) & ~3));  // [TK]
  uint64_t* bars = reinterpret_cast<uint64_t*>(klab + TK);               // S ready, O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, (uint32_t)TK);
    tmem_relinquish();
  }
  {
    const float* table = br == 0 ? p.table0 : p.table1;
    for (int i = threadIdx.x; i < tab_n; i += kTQ) tab[i] = table[(size_t)i * hpb + h] * kLog2e;  // softmax in base 2
  }

  // ---- staging: thread r owns key tokens r (and r + 128); its query token is one of them
  const T* src = reinterpret_cast<const T*>(p.src);
  const size_t hw = (size_t)p.H * p.W;
  const int ch0 = p.src_ch_off + (br * hpb + h) * kHP;
  const T* qg = src + ((size_t)n * p.src_planes + (ch0 >> 3)) * hw * 8;
  const T* kg = qg + (size_t)(p.qkv_stride >> 3) * hw * 8;
  const T* vg = kg + (size_t)(p.qkv_stride >> 3) * hw * 8;
  const uint32_t q_s = smem_u32(Qs), k_s = smem_u32(Ks), v_s = smem_u32(Vs);
  TokenPos qp = {-1, 0, 0, 0};
  int differs = 0;
  for (int kt = threadIdx.x, it = 0; kt < TK; kt += kTQ, ++it) {
    const int win = win0 + kt / N, t = kt - (kt / N) * N;
    const TokenPos tp = token_pos(p, Hs, Ws, nWx, nWin, win, t, sh, sw);
    if (p.shifted) {
      const TokenPos first = token_pos(p, Hs, Ws, nWx, nWin, win, 0, sh, sw);
      differs |= tp.lab != first.lab;
    }
    klab[kt] = (uint8_t)tp.lab;
    const bool is_q = kt == qoff + (int)threadIdx.x;
    if (is_q) qp = tp;
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
      const uint32_t ko = (uint32_t)(pl * TK + kt) * 16u, qo = (uint32_t)(pl * kTQ + (int)threadIdx.x) * 16u;
      if (tp.po >= 0) {
        const size_t g = ((size_t)pl * hw + tp.po) * 8;
        cp_async16(k_s + ko, kg + g);
        cp_async16(v_s + ko, vg + g);
        if (is_q) cp_async16(q_s + qo, qg + g);
      } else {
        sts_zero16(k_s + ko);
        sts_zero16(v_s + ko);
        if (is_q) sts_zero16(q_s + qo);
      }
    }
  }
  cp_async_wait_all();
  // dim 31 of every V row = 1: column 31 of O = sum of the rounded probabilities (head_dim < 32 always here)
  for (int kt = threadIdx.x; kt < TK; kt += kTQ) *reinterpret_cast<T*>(Vs + (size_t)(3 * TK + kt) * 16 + 14) = __float2bfloat16_rn(1.0f);
  fence_proxy_async_smem();  // generic-proxy writes (cp.async, st.shared) -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  const bool mixed = __syncthreads_or(differs) != 0;  // the shift mask only exists in windows that straddle the roll seam
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
  // query-side half of the bias address, label, while the MMA runs
  const int qpos = (qp.ty + Hs - 1) * tab_w + qp.tx + Ws - 1;
  const uint32_t qaddr = smem_u32(tab) + 4u * (uint32_t)qpos;
  const float scale2 = p.scale * kLog2e;
  // columns of this row's own window: 64-token windows share the tile in pairs (rows 0-63 | 64-127 <-> columns 0-63 | 64-127)
  const int blk = N < kTQ ? (int)threadIdx.x / N : 0;
  const int ncols = N < kTQ ? N : TK;
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t taddr = lane_base + (uint32_t)(blk * N);
  const uint8_t* klab_row = klab + blk * N;
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  float mx;
  if (Ws == 32) {
    mx = mixed ? softmax_pass1<32, true>(taddr, ncols, qaddr, scale2, klab_row, qp.lab) : softmax_pass1<32, false>(taddr, ncols, qaddr, scale2, klab_row, 0);
  } else if (Ws == 16) {
    mx = mixed ? softmax_pass1<16, true>(taddr, ncols, qaddr, scale2, klab_row, qp.lab) : softmax_pass1<16, false>(taddr, ncols, qaddr, scale2, klab_row, 0);
  } else {
    mx = mixed ? softmax_pass1<8, true>(taddr, ncols, qaddr, scale2, klab_row, qp.lab) : softmax_pass1<8, false>(taddr, ncols, qaddr, scale2, klab_row, 0);
  }
  // P: TK / 2 packed columns from column 0; this row's block at blk * N / 2, the other block (64-token windows) zeros
  softmax_pass2(taddr, lane_base + (uint32_t)(blk * N / 2), ncols, mx);
  if (N < kTQ) {
    const uint32_t other = lane_base + (uint32_t)((1 - blk) * N / 2);
    for (int c = 0; c < N / 2; c += 16) tmem_st16_zero(other + c);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();

  // ---- O = P V: A = P from TMEM (8 packed columns per K = 16 step), B = V [key][dim] MN-major; O in the dead score columns
  const uint32_t o_col = (uint32_t)TK / 2;
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(kTQ, kHP) | kIdescBMajorMN;
    for (int kk = 0; kk < TK / 16; ++kk) {
      // MN-major, no swizzle: 8 keys x 16 bytes (8 dims) = one core matrix; LBO = next 8 keys (128 B), SBO = next 8 dims (plane)
      const uint64_t db = make_smem_desc(v_s + (uint32_t)kk * 256u, 128u, (uint32_t)TK * 16u);
      umma_bf16_ts(tmem_base + o_col, tmem_base + (uint32_t)kk * 8u, db, idesc, kk);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  {
    uint32_t o[32];
    tmem_ld32(lane_base + o_col, o);
    tmem_ld_wait();
    if (qp.po >= 0) {
      const float inv = 1.0f / __uint_as_float(o[31]);
      T* dst = reinterpret_cast<T*>(p.dst);
      const int co = p.dst_ch_off + (br * hpb + h) * kHP;
      T* out = dst + (((size_t)n * p.dst_planes + (co >> 3)) * hw + qp.po) * 8;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)TK);
}

size_t winattn_tc_smem_bytes(int split_h, int split_w) {
  const int N = split_h * split_w, TK = std::max(N, kTQ);
  const int tab_n = (2 * split_h - 1) * (2 * split_w - 1);
  return (size_t)4 * kTQ * 16 + (size_t)8 * TK * 16 + (size_t)((tab_n + 3) & ~3) * 4 + TK + 2 * 8 + 16;
}

}  // namespace

bool winattn_tc_supported(int heads, int head_dim, int split_h, int split_w) {
  const int N = split_h * split_w;
  return heads % 2 == 0 && head_dim < kHP && (N == 64 || N == 128 || N == 256) && split_h % 8 == 0 && split_w % 8 == 0 && split_h <= 32 && split_w <= 32;
}

cudaError_t winattn_tc_configure() {
  return cudaFuncSetAttribute(winattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

cudaError_t launch_winattn_tc(const WinAttnParams& p, cudaStream_t s) {
  const int N = p.split_h * p.split_w;
  const int nWin = (p.Hp / p.split_h) * (p.Wp / p.split_w);
  const int tiles = N > kTQ ? nWin * (N / kTQ) : (nWin * N + kTQ - 1) / kTQ;
  const dim3 grid(tiles * p.n, p.heads / 2, 2);
  winattn_tc_kernel<<<grid, kTQ, winattn_tc_smem_bytes(p.split_h, p.split_w), s>>>(p);
  return cudaGetLastError();
}

}  // namespace rsb
