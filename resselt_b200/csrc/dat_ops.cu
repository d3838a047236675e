// Token-wise and attention kernels for DAT (Dual Aggregation Transformer) on planar-8 activations (token == pixel).
//
//   layernorm_kernel      LayerNorm over the channel dimension of every pixel (nn.LayerNorm call sites of
//                         /root/reference/resselt/archs/dat/arch.py:48,636,672,897,924)
//   dwconv3_kernel        depthwise 3x3 conv (+ folded BatchNorm + GELU, or * gate operand)   (arch.py:49,345,547)
//   winattn_kernel        fused shifted-window attention: one CTA per (window, head): K/V of the window in shared
//                         memory, one query per thread, QK^T + dynamic position bias + shift mask -> online softmax
//                         -> PV in registers; roll / pad / window (un)partition are pure addressing  (arch.py:224-267,456-482)
//   chanattn_*            channel attention: split-N Gram + norms reduction (deterministic two-stage), 30x30 softmax,
//                         per-token apply                                                        (arch.py:565-589)
//   aim_*                 adaptive interaction module: global average pool -> MLP -> channel map, per-pixel MLP ->
//                         spatial map, and the gated sum                                          (arch.py:492-508,594-607)
// All arithmetic is fp32 regardless of the storage type T (bf16 or fp32), so one set of kernels serves both plans.
#include <algorithm>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {
namespace {

template <typename T>
__device__ __forceinline__ float ld_ch(const T* base, int n, int planes, int plane0, int H, int W, int y, int x, int c) {
  return (float)base[planar_index(n, planes, plane0 + (c >> 3), H, W, y, x) + (c & 7)];
}
template <typename T>
__device__ __forceinline__ void st_ch(T* base, int n, int planes, int plane0, int H, int W, int y, int x, int c, float v) {
  base[planar_index(n, planes, plane0 + (c >> 3), H, W, y, x) + (c & 7)] = (T)v;
}
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float sigm_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------------ LayerNorm
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const __grid_constant__ TokenOpParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = (size_t)p.n * hw;
  const int C = p.channels, planes = (C + 7) >> 3;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / hw);
    const size_t pix = i - (size_t)n * hw;
    const T* s = src + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8 + pix * 8;
    float sum = 0.0f;
    for (int pl = 0; pl < planes; ++pl) {
      float v[8];
      load8<T>(s + (size_t)pl * hw * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pl * 8 + k < C) sum += v[k];
    }
    const float mean = sum / C;
    float sq = 0.0f;
    for (int pl = 0; pl < planes; ++pl) {
      float v[8];
      load8<T>(s + (size_t)pl * hw * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pl * 8 + k < C) sq += (v[k] - mean) * (v[k] - mean);
    }
    const float rstd = rsqrtf(sq / C + p.f0);
    T* d = dst + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + pix * 8;
    if (p.i0 >= 1) {  // statistics only (the normalisation is folded into the consuming conv: Epi::ln_stats)
      // i0 == 2: RMSNorm statistics {1 / (|x|_2 / sqrt(C) + eps), 0} (x / (rms + eps) has no mean term)
      *reinterpret_cast<float2*>(d) = p.i0 == 2 ? make_float2(1.0f / (sqrtf(sq / C + mean * mean) + p.f0), 0.0f) : make_float2(rstd, -mean * rstd);
      continue;
    }
    for (int pl = 0; pl < planes; ++pl) {
      float v[8];
      load8<T>(s + (size_t)pl * hw * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = pl * 8 + k;
        v[k] = c < C ? (v[k] - mean) * rstd * p.w0[c] + p.w1[c] : 0.0f;
      }
      store8<T>(d + (size_t)pl * hw * 8, v);
    }
  }
}

// bf16 storage, register-resident: a pixel's channels are split over the four quarter-warps (quarter q takes planes q, q + 4,
// ...; lanes 0..7 are 8 neighbouring pixels, so every load instruction covers 4 x 128 contiguous bytes), each thread
// issues all of its <= kPP 16-byte loads before the first use and the three LayerNorm passes (mean, centred variance,
// normalise) run out of registers: one HBM read + one write per element.  The generic kernel above re-reads the pixel three
// times through L1/L2.  Measured: all three formulations tried (three-pass, this one, and a cp.async.bulk-staged one with
// 184 KB per SM in flight) take 59-64 us per 180-channel 512^2 map (2.5 TB/s) — neither issue slots nor bytes in flight
// are what limits it; see DESIGN.md 3.5.
template <int kPP>
__global__ void __launch_bounds__(256, 2) layernorm_bf16_kernel(const __grid_constant__ TokenOpParams p) {
  using T = __nv_bfloat16;
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = (size_t)p.n * hw;
  const int C = p.channels, planes = (C + 7) >> 3;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  const int lane = threadIdx.x & 31, sub = lane >> 3;
  const size_t pix_in_block = (size_t)(threadIdx.x >> 5) * 8 + (lane & 7);  // 64 pixels per 256-thread block
  const float inv_c = 1.0f / (float)C;
  for (size_t i0 = (size_t)blockIdx.x * 64; i0 < total; i0 += (size_t)gridDim.x * 64) {
    const size_t i = i0 + pix_in_block;
    const bool live = i < total;
    const int n = live ? (int)(i / hw) : 0;
    const size_t pix = live ? i - (size_t)n * hw : 0;
    const T* s = src + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8 + pix * 8;
    uint4 raw[kPP];
#pragma unroll
    for (int j = 0; j < kPP; ++j) {
      const int pl = 4 * j + sub;
      raw[j] = (live && pl < planes) ? *reinterpret_cast<const uint4*>(s + (size_t)pl * hw * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    // fp32 copies, converted once (bf16 -> fp32 is a shift / mask); channels >= C of the last plane are forced to zero so
    // that none of the three passes below needs a per-element bound check
    float v[kPP][8];
#pragma unroll
    for (int j = 0; j < kPP; ++j) {
      const uint32_t w4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) v[j][2 * k] = __uint_as_float(w4[k] << 16), v[j][2 * k + 1] = __uint_as_float(w4[k] & 0xFFFF0000u);
      const int c0 = (4 * j + sub) * 8;
      if (c0 + 8 > C && c0 < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c0 + k >= C) v[j][k] = 0.0f;
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < kPP; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += v[j][k];
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    const float mean = sum * inv_c;
    // centred second moment: planes that do not exist (and the zeroed tail) would each add mean^2 — count them out exactly
    float sq = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kPP; ++j) {
      const int c0 = (4 * j + sub) * 8;
      if (c0 < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sq = fmaf(v[j][k] - mean, v[j][k] - mean, sq);
        cnt += c0 + 8 > C ? c0 + 8 - C : 0;
      }
    }
    sq -= (float)cnt * mean * mean;
    sq += __shfl_xor_sync(0xffffffffu, sq, 8);
    sq += __shfl_xor_sync(0xffffffffu, sq, 16);
    const float rstd = rsqrtf(fmaxf(sq, 0.0f) * inv_c + p.f0);
    const float shift = -mean * rstd;
    T* d = dst + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + pix * 8;
    if (p.i0 >= 1) {
      if (live && sub == 0)
        *reinterpret_cast<float2*>(d) = p.i0 == 2 ? make_float2(1.0f / (sqrtf(fmaxf(sq, 0.0f) * inv_c + mean * mean) + p.f0), 0.0f) : make_float2(rstd, shift);
      continue;
    }
#pragma unroll
    for (int j = 0; j < kPP; ++j) {
      const int pl = 4 * j + sub;
      if (!live || pl >= planes) continue;
      float g[8], b[8];
      if (pl * 8 + 8 <= C) {  // gamma / beta hold exactly C floats: vector loads only for whole planes
        const float4 g0 = *reinterpret_cast<const float4*>(p.w0 + pl * 8), g1 = *reinterpret_cast<const float4*>(p.w0 + pl * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(p.w1 + pl * 8), b1 = *reinterpret_cast<const float4*>(p.w1 + pl * 8 + 4);
        g[0] = g0.x, g[1] = g0.y, g[2] = g0.z, g[3] = g0.w, g[4] = g1.x, g[5] = g1.y, g[6] = g1.z, g[7] = g1.w;
        b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w, b[4] = b1.x, b[5] = b1.y, b[6] = b1.z, b[7] = b1.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = pl * 8 + k < C ? p.w0[pl * 8 + k] : 0.0f, b[k] = pl * 8 + k < C ? p.w1[pl * 8 + k] : 0.0f;
      }
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(fmaf(v[j][k], rstd, shift), g[k], b[k]);  // tail channels: g = b = 0 -> 0
      store8<T>(d + (size_t)pl * hw * 8, o);
    }
  }
}

// bf16 storage, streaming form: persistent CTAs (one per SM), a producer warp keeps a ring of kLnStages tiles in flight with
// 1-D bulk copies (cp.async.bulk, one 2 KB row of 128 pixels per 8-channel plane, mbarrier complete_tx), eight consumer warps
// normalise the tile that has landed.  Why: a streaming kernel needs bandwidth x latency bytes in flight (~70 KB per SM at
// 6.5 TB/s and ~1.7 us loaded latency); the register-resident kernel above holds 49 KB per SM only while its threads are in
// their load phase (measured 2.5 TB/s), whereas here the ring (4 x 46 KB for 180 channels) stays full while the consumers
// compute.  Consumer thread = (pixel, half): the two half-warps split the planes, statistics meet through one shuffle.
constexpr int kLnTile = 128;  // pixels per tile
constexpr int kLnConsumers = 256;
constexpr int kLnSplit = 2;  // consumer threads per pixel (lane groups of 32 / kLnSplit pixels, each taking every kLnSplit-th plane).  Four
                             // (16 consumer warps) measured slower: 32.9 vs 29.1 us per launch averaged over DAT's LayerNorms
__global__ void __launch_bounds__(kLnConsumers + 32, 1) layernorm_stream_kernel(const __grid_constant__ TokenOpParams p, int stages) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.channels, planes = (C + 7) >> 3;
  const int tiles_per_img = (int)((hw + kLnTile - 1) / kLnTile);
  const int total_tiles = p.n * tiles_per_img;
  const uint32_t stage_bytes = (uint32_t)planes * kLnTile * 16u;
  uint8_t* ring = ln_smem;
  float* gam = reinterpret_cast<float*>(ring + (size_t)stages * stage_bytes);  // [planes * 8]
  float* bet = gam + planes * 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(bet + planes * 8);
  uint64_t* empty = full + stages;
  for (int c = threadIdx.x; c < planes * 8; c += blockDim.x) gam[c] = c < C ? p.w0[c] : 0.0f, bet[c] = c < C ? p.w1[c] : 0.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], kLnConsumers / 32);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kLnConsumers / 32) {
    // ---- producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, t = tile - n * tiles_per_img;
        const size_t pix0 = (size_t)t * kLnTile;
        const uint32_t npix = (uint32_t)min((size_t)kLnTile, hw - pix0);
        const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8 + pix0 * 8;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&full[s], (uint32_t)planes * npix * 16u);
        for (int pl = 0; pl < planes; ++pl)
          ptx::bulk_load_1d(ring + (size_t)s * stage_bytes + (size_t)pl * kLnTile * 16, src + (size_t)pl * hw * 8, npix * 16u, &full[s]);
        if (++s == stages) s = 0, ph ^= 1;
      }
    }
    return;
  }
  // ---- consumers: warp w takes 32 / kLnSplit pixels of the tile; lane group `half` takes planes half, half + kLnSplit, ...
  constexpr int kPxW = 32 / kLnSplit;
  const int half = lane / kPxW, px = warp * kPxW + (lane % kPxW);
  const int full_planes = C >> 3, tail = C & 7;
  const float inv_c = 1.0f / (float)C;
  auto unpack = [](const uint4& r, float (&v)[8]) {
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) v[2 * k] = __uint_as_float(w4[k] << 16), v[2 * k + 1] = __uint_as_float(w4[k] & 0xFFFF0000u);
  };
  int s = 0;
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / tiles_per_img, t = tile - n * tiles_per_img;
    const size_t pix0 = (size_t)t * kLnTile;
    const int npix = (int)min((size_t)kLnTile, hw - pix0);
    ptx::mbar_wait(&full[s], ph);
    const uint4* mine = reinterpret_cast<const uint4*>(ring + (size_t)s * stage_bytes) + px;
    const bool live = px < npix;
    // One pass over the tile for the statistics: sums of (x - k) and (x - k)^2 with k = the first channel this thread sees
    // (a shift inside the data keeps the one-pass variance as accurate as the two-pass form), the lane groups of a pixel are merged with
    // the pairwise update of Chan et al.  The kernel is bound by instruction issue, not by HBM (ncu: 64 % issue-active on 9
    // warps per SM at 2.7 TB/s), so passes and per-element selects are what it pays for: the channel bound check is only made
    // on the one plane that can be partial.
    float k0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    if (live) {
      if (half < planes) {
        float v[8];
        unpack(mine[half * kLnTile], v);
        k0 = v[0];
      }
      for (int pl = half; pl < full_planes; pl += kLnSplit) {
        float v[8];
        unpack(mine[pl * kLnTile], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dlt = v[k] - k0;
          s1 += dlt;
          s2 = fmaf(dlt, dlt, s2);
        }
      }
      if (tail != 0 && (full_planes % kLnSplit) == half) {
        float v[8];
        unpack(mine[full_planes * kLnTile], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dlt = k < tail ? v[k] - k0 : 0.0f;
          s1 += dlt;
          s2 = fmaf(dlt, dlt, s2);
        }
      }
    }
    // channels held by this lane group; the (count, mean, M2) partials of a pixel are merged pairwise
    const int mine_planes = (full_planes + kLnSplit - 1 - half) / kLnSplit;
    float cnt = (float)(mine_planes * 8 + ((tail != 0 && (full_planes % kLnSplit) == half) ? tail : 0));
    float mean = cnt > 0.0f ? k0 + s1 / cnt : 0.0f;
    float m2 = cnt > 0.0f ? s2 - s1 * s1 / cnt : 0.0f;
#pragma unroll
    for (int sh = kPxW; sh <= 16; sh <<= 1) {
      const float cnt_o = __shfl_xor_sync(0xffffffffu, cnt, sh), mean_o = __shfl_xor_sync(0xffffffffu, mean, sh);
      const float m2_o = __shfl_xor_sync(0xffffffffu, m2, sh);
      const float tot = cnt + cnt_o, dm = mean_o - mean;
      const float inv = tot > 0.0f ? 1.0f / tot : 0.0f;
      // symmetric form: both partners must end up with bit-identical results
      mean = (cnt * mean + cnt_o * mean_o) * inv;
      m2 = m2 + m2_o + dm * dm * cnt * cnt_o * inv;
      cnt = tot;
    }
    const float rstd = rsqrtf(fmaxf(m2, 0.0f) * inv_c + p.f0);
    const float shift = -mean * rstd;
    if (p.i0 >= 1) {
      if (live && half == 0)
        *reinterpret_cast<float2*>(reinterpret_cast<T*>(p.dst) + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + (pix0 + px) * 8) =
            p.i0 == 2 ? make_float2(1.0f / (sqrtf(fmaxf(m2, 0.0f) * inv_c + mean * mean) + p.f0), 0.0f) : make_float2(rstd, shift);
    } else if (live) {
      T* d = reinterpret_cast<T*>(p.dst) + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + (pix0 + px) * 8;
      for (int pl = half; pl < planes; pl += kLnSplit) {
        float v[8], o[8];
        unpack(mine[pl * kLnTile], v);
        const float4 g0 = *reinterpret_cast<const float4*>(gam + pl * 8), g1 = *reinterpret_cast<const float4*>(gam + pl * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bet + pl * 8), b1 = *reinterpret_cast<const float4*>(bet + pl * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(fmaf(v[k], rstd, shift), g[k], b[k]);  // tail channels: gamma = beta = 0 -> 0
        store8<T>(d + (size_t)pl * hw * 8, o);
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);  // this warp is done reading the stage
    if (++s == stages) s = 0, ph ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
template <typename T>
__global__ void __launch_bounds__(256) dwconv3_kernel(const __grid_constant__ TokenOpParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.channels, planes = (C + 7) >> 3;
  const size_t total = (size_t)p.n * planes * hw;
  const T* src = reinterpret_cast<const T*>(p.src);
  const T* mul = reinterpret_cast<const T*>(p.src2);
  T* dst = reinterpret_cast<T*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.W);
    const int y = (int)((i / p.W) % p.H);
    const int pl = (int)((i / hw) % planes);
    const int n = (int)(i / (hw * planes));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = pl * 8 + k < C ? p.w1[pl * 8 + k] : 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = y + ky - 1;
      if (sy < 0 || sy >= p.H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = x + kx - 1;
        if (sx < 0 || sx >= p.W) continue;
        float v[8];
        load8<T>(src + planar_index(n, p.src_planes, p.src_plane0 + pl, p.H, p.W, sy, sx), v);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (pl * 8 + k < C) acc[k] = fmaf(v[k], p.w0[(pl * 8 + k) * 9 + ky * 3 + kx], acc[k]);
      }
    }
    if (p.i0 == RSB_ACT_GELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = gelu_f(acc[k]);
    }
    if (mul != nullptr) {
      float g[8];
      load8<T>(mul + planar_index(n, p.src2_planes, p.src2_plane0 + pl, p.H, p.W, y, x), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= g[k];
    }
    store8<T>(dst + planar_index(n, p.dst_planes, p.dst_plane0 + pl, p.H, p.W, y, x), acc);
  }
}

// bf16 storage: each thread owns one pixel column of one 8-channel plane over a vertical run of rows.  The plane's 72
// weights live in registers for the whole run and the 3 x 3 window of 16-byte pixel chunks slides down, so an output costs
// three loads (two of them L1 hits) instead of nine plus 72 weight loads; GELU uses the MUFU erf approximation of the
// conv epilogues (|err| <= 1.5e-7).  First version: 170 us per launch on DAT 4x 512^2.
constexpr int kDwRows = 8;  // rows per thread (48 rows, to amortise the 80 weight loads of the prologue, measured 151 instead of 104 us
                            // per launch: the kernel lives on the number of independent row chains in flight, not on instruction count)
// (a0, a1) += (v0, v1) * (w0, w1) as one packed fp32 FMA (sm_100 FFMA2): the depthwise kernel is bound by instruction issue
__device__ __forceinline__ void fma2_pair(float& a0, float& a1, float v0, float v1, float w0, float w1) {
  asm("{\n\t"
      ".reg .b64 rv, rw, ra;\n\t"
      "mov.b64 rv, {%2, %3};\n\t"
      "mov.b64 rw, {%4, %5};\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "fma.rn.f32x2 ra, rv, rw, ra;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t"
      "}"
      : "+f"(a0), "+f"(a1)
      : "f"(v0), "f"(v1), "f"(w0), "f"(w1));
}

__global__ void __launch_bounds__(128) dwconv3_bf16_kernel(const __grid_constant__ TokenOpParams p) {
  using T = __nv_bfloat16;
  const int C = p.channels, planes = (C + 7) >> 3;
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int yb = blockIdx.y * kDwRows;
  const int pl = blockIdx.z % planes, n = blockIdx.z / planes;
  if (x >= p.W) return;
  const T* src = reinterpret_cast<const T*>(p.src) + planar_index(n, p.src_planes, p.src_plane0 + pl, p.H, p.W, 0, 0);
  const T* mul = reinterpret_cast<const T*>(p.src2);
  T* dst = reinterpret_cast<T*>(p.dst);
  float w[9][8], bias[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = pl * 8 + k;
    bias[k] = c < C ? p.w1[c] : 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t][k] = c < C ? p.w0[c * 9 + t] : 0.0f;
  }
  const uint4 zero = make_uint4(0, 0, 0, 0);
  auto ldrow = [&](int y, uint4 (&o)[3]) {
    if (y < 0 || y >= p.H) {
      o[0] = o[1] = o[2] = zero;
      return;
    }
    const uint4* r = reinterpret_cast<const uint4*>(src + ((size_t)y * p.W + x) * 8);
    o[0] = x > 0 ? r[-1] : zero;
    o[1] = r[0];
    o[2] = x + 1 < p.W ? r[1] : zero;
  };
  // ring of four rows: the load of row y + 2 is issued while row y is computed — two row loads in flight per thread (with one,
  // 24 resident warps x 512 B per SM could not cover HBM latency: 1.4 TB/s)
  uint4 win[4][3];
  ldrow(yb - 1, win[0]);
  ldrow(yb, win[1]);
  ldrow(yb + 1, win[2]);
#pragma unroll
  for (int r = 0; r < kDwRows; ++r) {
    const int y = yb + r;
    if (y < p.H) {
      ldrow(y + 2, win[(r + 3) % 4]);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = bias[k];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float v[8];
          unpack8<T>(win[(r + ky) % 4][kx], v);
#pragma unroll
          for (int k = 0; k < 8; k += 2) fma2_pair(acc[k], acc[k + 1], v[k], v[k + 1], w[ky * 3 + kx][k], w[ky * 3 + kx][k + 1]);
        }
      if (p.i0 == RSB_ACT_GELU) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = activate<true, RSB_ACT_GELU>(RSB_ACT_GELU, acc[k], 0.0f, 0.0f);
      }
      if (mul != nullptr) {
        float g[8];
        load8<T>(mul + planar_index(n, p.src2_planes, p.src2_plane0 + pl, p.H, p.W, y, x), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= g[k];
      }
      store8<T>(dst + planar_index(n, p.dst_planes, p.dst_plane0 + pl, p.H, p.W, y, x), acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------ window attention
constexpr int kHD = 32;  // head_dim padded (<= 32 supported)

template <typename T>
__global__ void __launch_bounds__(256) winattn_kernel(const __grid_constant__ WinAttnParams p) {
  extern __shared__ float sm[];
  const int br = blockIdx.z, h = blockIdx.y;
  const int Hs = br == 0 ? p.split_h : p.split_w, Ws = br == 0 ? p.split_w : p.split_h;
  const int N = Hs * Ws;
  const int sh = p.shifted ? Hs / 2 : 0, sw = p.shifted ? Ws / 2 : 0;
  const int nWx = p.Wp / Ws, nWy = p.Hp / Hs;
  const int win = blockIdx.x % (nWx * nWy), n = blockIdx.x / (nWx * nWy);
  const int wy = win / nWx, wx = win - wy * nWx;
  const int d = p.head_dim;
  const int tab_w = 2 * Ws - 1, tab_n = (2 * Hs - 1) * tab_w;
  float* Ks = sm;                       // [N][kHD]
  float* Vs = Ks + (size_t)N * kHD;     // [N][kHD]
  float* tab = Vs + (size_t)N * kHD;    // [tab_n] bias of this head
  int* labs = reinterpret_cast<int*>(tab + tab_n);  // [N]
  const float* table = br == 0 ? p.table0 : p.table1;
  const int hpb = p.heads / 2;  // heads per branch
  for (int i = threadIdx.x; i < tab_n; i += blockDim.x) tab[i] = table[(size_t)i * hpb + h];

  const int t = threadIdx.x;
  const bool active = t < N;
  const int ty = active ? t / Ws : 0, tx = active ? t - ty * Ws : 0;
  const int yr = wy * Hs + ty, xr = wx * Ws + tx;       // coordinates in the rolled, padded image
  const int yo = (yr + sh) % p.Hp, xo = (xr + sw) % p.Wp;  // where that token lives in the un-rolled image
  const bool inb = active && yo < p.H && xo < p.W;      // padded tokens have q = k = v = 0
  const T* src = reinterpret_cast<const T*>(p.src);
  const int half = p.dim / 2;
  const int cq = p.src_ch_off + br * half + h * d;
  float q[kHD];
#pragma unroll
  for (int c = 0; c < kHD; ++c) {
    float qv = 0.0f, kv = 0.0f, vv = 0.0f;
    if (inb && c < d) {
      qv = ld_ch<T>(src, n, p.src_planes, 0, p.H, p.W, yo, xo, cq + c);
      kv = ld_ch<T>(src, n, p.src_planes, 0, p.H, p.W, yo, xo, cq + p.qkv_stride + c);
      vv = ld_ch<T>(src, n, p.src_planes, 0, p.H, p.W, yo, xo, cq + 2 * p.qkv_stride + c);
    }
    q[c] = qv * p.scale;
    if (active) Ks[t * kHD + c] = kv, Vs[t * kHD + c] = vv;
  }
  if (active) {
    int lab = 0;
    if (p.shifted) {
      const int ry = yr < p.Hp - Hs ? 0 : (yr < p.Hp - sh ? 1 : 2);
      const int rx = xr < p.Wp - Ws ? 0 : (xr < p.Wp - sw ? 1 : 2);
      lab = 3 * ry + rx;
    }
    labs[t] = lab;
  }
  __syncthreads();
  if (!active) return;
  const int my_lab = labs[t];
  float m = -INFINITY, l = 0.0f;
  float o[kHD];
#pragma unroll
  for (int c = 0; c < kHD; ++c) o[c] = 0.0f;
  int jy = 0, jx = 0;
  for (int j = 0; j < N; ++j) {
    const float4* k4 = reinterpret_cast<const float4*>(Ks + j * kHD);
    float s = 0.0f;
#pragma unroll
    for (int c4 = 0; c4 < kHD / 4; ++c4) {
      const float4 kk = k4[c4];
      s = fmaf(q[4 * c4 + 0], kk.x, s);
      s = fmaf(q[4 * c4 + 1], kk.y, s);
      s = fmaf(q[4 * c4 + 2], kk.z, s);
      s = fmaf(q[4 * c4 + 3], kk.w, s);
    }
    s += tab[(ty - jy + Hs - 1) * tab_w + (tx - jx + Ws - 1)];
    if (p.shifted && labs[j] != my_lab) s += -100.0f;
    if (s > m) {  // rescale the running sums to the new maximum
      const float corr = __expf(m - s);
      l *= corr;
#pragma unroll
      for (int c = 0; c < kHD; ++c) o[c] *= corr;
      m = s;
    }
    const float pj = __expf(s - m);
    l += pj;
    const float4* v4 = reinterpret_cast<const float4*>(Vs + j * kHD);
#pragma unroll
    for (int c4 = 0; c4 < kHD / 4; ++c4) {
      const float4 vv = v4[c4];
      o[4 * c4 + 0] = fmaf(pj, vv.x, o[4 * c4 + 0]);
      o[4 * c4 + 1] = fmaf(pj, vv.y, o[4 * c4 + 1]);
      o[4 * c4 + 2] = fmaf(pj, vv.z, o[4 * c4 + 2]);
      o[4 * c4 + 3] = fmaf(pj, vv.w, o[4 * c4 + 3]);
    }
    if (++jx == Ws) jx = 0, ++jy;
  }
  if (inb) {
    const float inv = 1.0f / l;
    T* dst = reinterpret_cast<T*>(p.dst);
    const int co = p.dst_ch_off + br * half + h * d;
#pragma unroll
    for (int c = 0; c < kHD; ++c)
      if (c < d) st_ch<T>(dst, n, p.dst_planes, 0, p.H, p.W, yo, xo, co + c, o[c] * inv);
  }
}

// ---- tensor-core version (bf16 storage): one CTA per (window, head), one warp per 16 queries, FlashAttention-2 style.
// S = Q K^T and O = P V run on warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) over 64-key chunks with an online
// softmax in registers; the score scale, position bias, shift mask and softmax are fp32, P is rounded to bf16 for the PV
// product (as in every bf16 attention kernel).  Q, K and V rows sit in shared memory as bf16 [token][32 dims + 8] (80-byte rows:
// 16-byte aligned and conflict-free for ldmatrix).
//
// The kernel is bound by instruction issue, not by the tensor pipe or memory (ncu, round 1: issue slots 61-70 % busy at 24 %
// occupancy, tensor pipe ~20 %, DRAM ~10 %), so the work per score is what counts.  The first MMA version spent ~19 instructions
// per (query, key) pair; this one ~8:
//   * K and V fragments come from ldmatrix.x4 (V through .trans, so V is staged like K instead of transposed element by element):
//     16 loads per 64-key chunk instead of 64;
//   * the relative-position bias of the two adjacent keys a thread holds per accumulator tile (same window row when the window
//     width is even, so their table entries are neighbours) is ONE 8-byte load: the table is kept three times at a stride that
//     makes [base + (a + b) * D + 4 i] 8-byte aligned whatever the parities a, b of the query and key halves of the index i
//     (copy 0 and 2 serve even i, copy 1 odd i), so the address is a single subtraction of a per-key word from a per-row word;
//   * the shift mask is skipped for windows whose tokens all carry one label (all but the last row / column of windows).
// A tcgen05 formulation (S in TMEM) does not change this bound: at head_dim 30 the MMAs are ~1/8 of the instructions, and the
// softmax (max, subtract, ex2, convert: 4.5 instructions per score on 128 lanes) is the floor either way.
constexpr int kWaQS = kHD + 8;  // Q / K / V row stride in bf16 elements

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) { ldsm_x4_trans(r, ptx::smem_u32(smem_row)); }
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ptx::smem_u32(smem_row)));
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t smem_addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_addr));
  return v;
}
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// distance (in floats) between the copies of the bias table: odd, and at least the table's length
__host__ __device__ inline int winattn_tab_stride(int tab_n) { return ((tab_n + 2) & ~1) - 1; }

size_t winattn_mma_smem_bytes(int split_h, int split_w) {
  const int N = split_h * split_w, NK = (N + 63) / 64 * 64, NQ = (N + 15) / 16 * 16;
  const int tab_n = (2 * split_h - 1) * (2 * split_w - 1);
  return (size_t)(NQ + 2 * NK) * kWaQS * 2 + (size_t)(2 * winattn_tab_stride(tab_n) + tab_n + 1) * 4 + (size_t)NK * 12;
}

// kThreads = 128 serves windows of <= 64 tokens (SwinIR's 8x8): a tighter register budget keeps 5 CTAs per SM resident
// (the kernel's phases — stage, attend, store — are latency-bound chains; residency is what overlaps them).
template <int kThreads, int kMinBlocks, bool kLoop>
__global__ void __launch_bounds__(kThreads, kMinBlocks) winattn_mma_kernel(const __grid_constant__ WinAttnParams p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  using T = __nv_bfloat16;
  const int br = blockIdx.z, h = blockIdx.y;
  const int Hs = br == 0 ? p.split_h : p.split_w, Ws = br == 0 ? p.split_w : p.split_h;
  const int N = Hs * Ws, NK = (N + 63) / 64 * 64, NQ = (N + 15) / 16 * 16;
  const int sh = p.shifted ? Hs / 2 : 0, sw = p.shifted ? Ws / 2 : 0;
  const int nWx = p.Wp / Ws, nWy = p.Hp / Hs;
  const int win = blockIdx.x % (nWx * nWy), n = blockIdx.x / (nWx * nWy);
  const int wy = win / nWx, wx = win - wy * nWx;
  const int d = p.head_dim;
  const int tab_w = 2 * Ws - 1, tab_n = (2 * Hs - 1) * tab_w;
  const int R1 = winattn_tab_stride(tab_n);
  T* Qs = reinterpret_cast<T*>(smraw);            // [NQ][kWaQS]
  T* Ks = Qs + (size_t)NQ * kWaQS;                // [NK][kWaQS]
  T* Vs = Ks + (size_t)NK * kWaQS;                // [NK][kWaQS]
  float* tab3 = reinterpret_cast<float*>(Vs + (size_t)NK * kWaQS);  // bias of this head (x log2 e), copies at 0, R1, 2 R1
  int* kval = reinterpret_cast<int*>(tab3 + 2 * R1 + tab_n + 1);    // [NK] 4 * kpos - (kpos & 1) * 4 R1   (kpos = jy * tab_w + jx)
  int* kmeta = kval + NK;                                           // [NK] 4 * kpos | label << 20 | (not a key) << 30
  int* tokoff = kmeta + NK;                                         // [NK] pixel index y * W + x of the token, -1: padding
  const float* table = br == 0 ? p.table0 : p.table1;
  const int hpb = p.heads / 2;
  for (int i = threadIdx.x; i < tab_n; i += blockDim.x) {
    const float v = table[(size_t)i * hpb + h] * kLog2e;  // softmax in base 2
    tab3[i] = v, tab3[R1 + i] = v, tab3[2 * R1 + i] = v;
  }
  const float scale2 = p.scale * kLog2e;
  const bool ones_col = d < kHD;  // spare V column carries the softmax denominators (see below)

  const T* src = reinterpret_cast<const T*>(p.src);
  const int half = p.dim / 2;
  const int cq = p.src_ch_off + br * half + h * d;
  // Staging.  The tiles are zero-filled first (padded tokens, head dims d..31 and chunk padding must read as 0), then every
  // (matrix, 8-channel plane, token) item is ONE 16-byte load — consecutive threads take consecutive tokens, i.e. consecutive
  // 16-byte chunks of a plane row — whose channels inside [cq, cq + d) go to the token's row.  A head's 30 channels start at
  // any channel offset, so it touches up to 5 planes per matrix.
  {
    uint4* z = reinterpret_cast<uint4*>(smraw);
    const int nz = (NQ + 2 * NK) * kWaQS * 2 / 16;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  int lab0 = 0;  // label of the window's first token
  if (p.shifted) {
    const int yr = wy * Hs, xr = wx * Ws;
    lab0 = 3 * (yr < p.Hp - Hs ? 0 : (yr < p.Hp - sh ? 1 : 2)) + (xr < p.Wp - Ws ? 0 : (xr < p.Wp - sw ? 1 : 2));
  }
  int differs = 0;
  for (int t = threadIdx.x; t < NK; t += blockDim.x) {
    const bool active = t < N;
    const int ty = active ? t / Ws : 0, tx = active ? t - ty * Ws : 0;
    const int yr = wy * Hs + ty, xr = wx * Ws + tx;  // coordinates in the rolled, padded image
    int yo = yr + sh, xo = xr + sw;                  // where that token lives in the un-rolled image
    if (yo >= p.Hp) yo -= p.Hp;
    if (xo >= p.Wp) xo -= p.Wp;
    int lab = lab0;
    if (p.shifted && active) {
      const int ry = yr < p.Hp - Hs ? 0 : (yr < p.Hp - sh ? 1 : 2);
      const int rx = xr < p.Wp - Ws ? 0 : (xr < p.Wp - sw ? 1 : 2);
      lab = 3 * ry + rx;
    }
    differs |= lab != lab0;
    const int kpos = ty * tab_w + tx;
    kval[t] = 4 * kpos - (kpos & 1) * 4 * R1;
    kmeta[t] = (4 * kpos) | (lab << 20) | (active ? 0 : 1 << 30);
    tokoff[t] = (active && yo < p.H && xo < p.W) ? yo * p.W + xo : -1;  // padded tokens have q = k = v = 0
  }
  const bool mixed = __syncthreads_or(differs) != 0;  // the shift mask only exists in windows that straddle the roll seam
  if (ones_col)
    for (int t = threadIdx.x; t < NK; t += blockDim.x) Vs[t * kWaQS + kHD - 1] = __float2bfloat16_rn(1.0f);
  if constexpr (!kLoop) {
    // small windows (several threads per token): the items are walked with a run-time stride.  A/B on one box, SwinIR 8x8
    // windows: 226 us per launch against 245 us with the unrolled walk below (which wins on 256-token windows: 371 against 384)
    constexpr int kPl = (kHD + 7 + 7) / 8;  // planes a head can straddle
    const size_t ps = (size_t)p.H * p.W * 8;
    const bool even = ((cq | p.qkv_stride | d) & 1) == 0;
    const int groups = max(1, (int)blockDim.x / N), g = threadIdx.x / N;
    const int tstep = (int)blockDim.x >= N ? N : (int)blockDim.x;
    for (int t = threadIdx.x - g * N; t < N && g < groups; t += tstep) {
      const int po = tokoff[t];
      if (po < 0) continue;
      const T* tok = src + ((size_t)n * p.src_planes * p.H * p.W + po) * 8;
      constexpr int kItems = (3 * kPl + 1) / 2;
      for (int mp0 = g; mp0 < 3 * kPl; mp0 += groups * kItems) {
        uint4 v[kItems];
#pragma unroll
        for (int it = 0; it < kItems; ++it) {
          const int mp = mp0 + it * groups;
          const int m = mp / kPl, pl = mp - m * kPl;
          const int cm = cq + m * p.qkv_stride;
          const int plane = (cm >> 3) + pl;
          if (mp < 3 * kPl && plane * 8 - cm < d) v[it] = *reinterpret_cast<const uint4*>(tok + (size_t)plane * ps);
        }
#pragma unroll
        for (int it = 0; it < kItems; ++it) {
          const int mp = mp0 + it * groups;
          const int m = mp / kPl, pl = mp - m * kPl;
          const int cm = cq + m * p.qkv_stride;
          const int cbase = ((cm >> 3) + pl) * 8 - cm;
          if (mp >= 3 * kPl || cbase >= d) continue;
          T* out = (m == 0 ? Qs : (m == 1 ? Ks : Vs)) + t * kWaQS;
          if (even) {
            const uint32_t* e = reinterpret_cast<const uint32_t*>(&v[it]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((unsigned)(cbase + 2 * k) < (unsigned)d) *reinterpret_cast<uint32_t*>(out + cbase + 2 * k) = e[k];
          } else {
            const T* e = reinterpret_cast<const T*>(&v[it]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((unsigned)(cbase + k) < (unsigned)d) out[cbase + k] = e[k];
          }
        }
      }
    }
  } else {
    // thread = (token, share g of the items).  The (matrix m, plane pl) items are walked by fully unrolled loops, so everything
    // that does not depend on the token — which plane, which of its 8 channels belong to the head, where they go — is uniform
    // arithmetic done once; an item costs its thread one 16-byte load and up to four 4-byte stores.  (Walking the items with a
    // run-time stride cost ~35 instructions per item, a third of the kernel's instructions on 8x32 windows and two thirds on 8x8.)
    constexpr int kPl = (kHD + 7 + 7) / 8;  // planes a head can straddle
    const size_t ps = (size_t)p.H * p.W * 8;
    const bool even = ((cq | p.qkv_stride | d) & 1) == 0;  // channel pairs never straddle the head: 4-byte stores
    // blocks with at least 2N / 4N threads: 2 / 4 threads share a token (items split by index mod 2 / 4); smaller blocks: each
    // thread walks several tokens
    const int gbits = (int)blockDim.x >= 4 * N ? 2 : ((int)blockDim.x >= 2 * N ? 1 : 0);
    const int g = threadIdx.x / N;
    const int tstep = (int)blockDim.x >= N ? N : (int)blockDim.x;
    int r8[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) r8[m] = (cq + m * p.qkv_stride) & 7;  // first channel of the head inside its first plane
    for (int t = threadIdx.x - g * N; t < N && g < (1 << gbits); t += tstep) {
      const int po = tokoff[t];
      if (po < 0) continue;
      const T* tok = src + ((size_t)n * p.src_planes * p.H * p.W + po) * 8;
      const T* tokm[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) tokm[m] = tok + (size_t)((cq + m * p.qkv_stride) >> 3) * ps;
      // all of a thread's loads are issued before the first value is used: ONE DRAM round trip per CTA (two batches measured
      // 248 instead of 226 us per launch on SwinIR's 8x8 windows: the kernel's phases are latency chains)
      constexpr int kBatch = 3 * kPl;
#pragma unroll
      for (int b0 = 0; b0 < 3 * kPl; b0 += kBatch) {
        uint4 v[kBatch];
#pragma unroll
        for (int it = 0; it < kBatch; ++it) {
          const int mp = b0 + it, m = mp / kPl, pl = mp - m * kPl;
          if (mp < 3 * kPl && (mp & ((1 << gbits) - 1)) == g && pl * 8 - r8[m] < d) v[it] = *reinterpret_cast<const uint4*>(tokm[m] + (size_t)pl * ps);
        }
#pragma unroll
        for (int it = 0; it < kBatch; ++it) {
          const int mp = b0 + it, m = mp / kPl, pl = mp - m * kPl;
          if (mp >= 3 * kPl) continue;
          const int cbase = pl * 8 - r8[m];  // head dim of the plane's first channel (negative: the previous head's)
          if ((mp & ((1 << gbits) - 1)) != g || cbase >= d) continue;
          T* out = (m == 0 ? Qs : (m == 1 ? Ks : Vs)) + t * kWaQS;
          if (even) {
            const uint32_t* e = reinterpret_cast<const uint32_t*>(&v[it]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if ((unsigned)(cbase + 2 * k) < (unsigned)d) *reinterpret_cast<uint32_t*>(out + cbase + 2 * k) = e[k];
          } else {
            const T* e = reinterpret_cast<const T*>(&v[it]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((unsigned)(cbase + k) < (unsigned)d) out[cbase + k] = e[k];
          }
        }
      }
    }  // tokens of this thread
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 2, t4 = lane & 3;
  T* dst = reinterpret_cast<T*>(p.dst);
  const int co = p.dst_ch_off + br * half + h * d;
  const bool pairs = (Ws & 1) == 0;  // keys 2k, 2k+1 share a window row: their bias entries are neighbours
  const uint32_t tab_s = ptx::smem_u32(tab3);
  // ldmatrix row addresses of this lane: K tile rows (lane & 7), 8-dim block (lane >> 3); V: key (lane & 15), dim block (lane >> 4)
  const uint32_t k_ld = ptx::smem_u32(Ks) + (uint32_t)(((lane & 7) * kWaQS + (lane >> 3) * 8) * 2);
  const uint32_t v_ld = ptx::smem_u32(Vs) + (uint32_t)(((lane & 15) * kWaQS + (lane >> 4) * 8) * 2);
  // a warp takes 16 queries at a time; with 256-token windows a 256-thread CTA walks two query tiles per warp, so that two
  // CTAs fit an SM and one window's staging overlaps the other's attention (one 512-thread CTA per SM left the SM idle
  // during every staging phase)
  int r0 = warp * 16;
  if (r0 >= NQ) return;
  do {  // (single pass, known at compile time, for the 128-thread variant)
  const int qa = r0 + g4, qb = qa + 8;  // the two query rows this thread holds fragments of
  // query-side halves of the bias address and the shift-mask labels
  uint32_t qaddr[2];
  int qlab[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int t = e == 0 ? qa : qb;
    const int tt = t < N ? t : 0;
    const int ty = tt / Ws, tx = tt - ty * Ws;
    const int qpos = (ty + Hs - 1) * tab_w + tx + Ws - 1;
    // pairs: entries (qpos - kpos - 1, qpos - kpos) of the copy that makes the pair 8-byte aligned; else entry qpos - kpos of copy 0
    qaddr[e] = pairs ? tab_s + (uint32_t)(((qpos - 1) & 1) * 4 * R1 + 4 * (qpos - 1)) : tab_s + (uint32_t)(4 * qpos);
    qlab[e] = (kmeta[tt] >> 20) & 0xF;
  }
  uint32_t qf[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const T* qr = Qs + ks * 16 + t4 * 2;
    qf[ks][0] = *reinterpret_cast<const uint32_t*>(qr + (size_t)qa * kWaQS);
    qf[ks][1] = *reinterpret_cast<const uint32_t*>(qr + (size_t)qb * kWaQS);
    qf[ks][2] = *reinterpret_cast<const uint32_t*>(qr + (size_t)qa * kWaQS + 8);
    qf[ks][3] = *reinterpret_cast<const uint32_t*>(qr + (size_t)qb * kWaQS + 8);
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.0f, 0.0f};
  float o[4][4];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.0f;

  for (int kc = 0; kc < NK; kc += 64) {
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[nt][e] = 0.0f;
      uint32_t kf[4];
      ldsm_x4(kf, k_ld + (uint32_t)((kc + nt * 8) * kWaQS * 2));
      mma_bf16_16816(sc[nt], qf[0], kf[0], kf[1]);
      mma_bf16_16816(sc[nt], qf[1], kf[2], kf[3]);
    }
    // scale + position bias (both pre-multiplied by log2 e: the softmax runs on ex2), then — only where they exist — the
    // shift mask and the chunk padding, as warp-uniform branches around their own loops; chunk maxima last
    if (pairs) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int kv = kval[kc + nt * 8 + t4 * 2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float2 b = lds_f32x2(qaddr[r] - (uint32_t)kv);  // .y: this key, .x: the next one
          sc[nt][2 * r] = fmaf(sc[nt][2 * r], scale2, b.y);
          sc[nt][2 * r + 1] = fmaf(sc[nt][2 * r + 1], scale2, b.x);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ki = kmeta[kc + nt * 8 + t4 * 2 + (e & 1)];
          float b;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b) : "r"(qaddr[e >> 1] - (uint32_t)(ki & 0xFFFFF)));
          sc[nt][e] = fmaf(sc[nt][e], scale2, b);
        }
    }
    if (mixed) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ki = kmeta[kc + nt * 8 + t4 * 2 + (e & 1)];
          if (((ki >> 20) & 0xF) != qlab[e >> 1]) sc[nt][e] += -100.0f * kLog2e;
        }
    }
    if (NK != N) {  // keys beyond the window (chunk padding)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kmeta[kc + nt * 8 + t4 * 2 + (e & 1)] >> 30) sc[nt][e] = -INFINITY;
    }
    float cm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      cm[0] = fmaxf(cm[0], fmaxf(sc[nt][0], sc[nt][1]));
      cm[1] = fmaxf(cm[1], fmaxf(sc[nt][2], sc[nt][3]));
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 1));
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 2));
      const float mn = fmaxf(m[r], cm[r]);
      corr[r] = ex2_approx(m[r] - mn);  // first chunk: 2^(-inf) = 0
      m[r] = mn;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      o[dt][0] *= corr[0], o[dt][1] *= corr[0];
      o[dt][2] *= corr[1], o[dt][3] *= corr[1];
    }
    // P = 2^(S - m) as bf16 A fragments.  O / l must be a true weighted mean of the ROUNDED probabilities: with head_dim < 32 the
    // row sums come for free from the PV product itself (column kHD-1 of V is all ones, so column kHD-1 of O accumulates
    // sum_j P_ij); only head_dim == 32 adds the rounded values up by hand.
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ex2_approx(sc[nt][0] - m[0]), p1 = ex2_approx(sc[nt][1] - m[0]);
      const float p2 = ex2_approx(sc[nt][2] - m[1]), p3 = ex2_approx(sc[nt][3] - m[1]);
      const uint32_t lo = pack_bf16x2(p0, p1), hi = pack_bf16x2(p2, p3);
      if (!ones_col) {
        l[0] += __uint_as_float(lo << 16) + __uint_as_float(lo & 0xFFFF0000u);
        l[1] += __uint_as_float(hi << 16) + __uint_as_float(hi & 0xFFFF0000u);
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = lo;
      pf[nt >> 1][(nt & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        // matrices of one ldmatrix.x4.trans: keys [0, 8) and [8, 16) of the 16-key step for dim block 2 dp, then for 2 dp + 1
        uint32_t vf[4];
        ldsm_x4_trans(vf, v_ld + (uint32_t)(((kc + kt * 16) * kWaQS + dp * 16) * 2));
        mma_bf16_16816(o[2 * dp], pf[kt], vf[0], vf[1]);
        mma_bf16_16816(o[2 * dp + 1], pf[kt], vf[2], vf[3]);
      }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (ones_col) {
      l[r] = __shfl_sync(0xffffffffu, o[3][2 * r + 1], lane | 3);  // column kHD-1 = 31 lives in the t4 == 3 lane of each row group
    } else {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = r == 0 ? qa : qb;
    const int po = t < N ? tokoff[t] : -1;
    if (po < 0) continue;
    const float inv = 1.0f / l[r];
    const size_t base = ((size_t)n * p.dst_planes * p.H * p.W + po) * 8;
    const size_t ps = (size_t)p.H * p.W * 8;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int c = dt * 8 + t4 * 2;  // dims c, c + 1
      const int ch = co + c;
      if (c + 1 < d && (ch & 1) == 0) {
        *reinterpret_cast<uint32_t*>(dst + base + (size_t)(ch >> 3) * ps + (ch & 7)) = pack_bf16x2(o[dt][2 * r] * inv, o[dt][2 * r + 1] * inv);
      } else {
        if (c < d) dst[base + (size_t)(ch >> 3) * ps + (ch & 7)] = __float2bfloat16_rn(o[dt][2 * r] * inv);
        if (c + 1 < d) dst[base + (size_t)((ch + 1) >> 3) * ps + ((ch + 1) & 7)] = __float2bfloat16_rn(o[dt][2 * r + 1] * inv);
      }
    }
  }
  r0 += kThreads / 2;
  } while (kLoop && r0 < NQ);  // query tiles
}

// ------------------------------------------------------------------------------------------------ channel attention
constexpr int kTok = 64;  // tokens per shared-memory tile of the Gram reduction

// partial[n][head][block][d*d + 2d]: Gram q^T k and squared column norms of q and k over this block's token range.
// Register-tiled: the 32 x 32 (head_dim padded) Gram matrix is cut into 8 x 8 tiles of 4 x 4 accumulators; the 256
// threads are 4 token groups x 64 tiles, so one token costs a thread two 16-byte shared-memory loads for 16 FMAs (the
// first version read two scalars per FMA and was shared-memory bound: 769 us per launch on DAT 4x 512^2).  Loads are
// whole 16-byte pixel chunks of the planes the head's channels live in (coalesced along tokens).  Summation order is
// fixed (token groups are reduced in order), so results are run-to-run deterministic.
template <typename T>
__global__ void __launch_bounds__(256) chanattn_reduce_kernel(const __grid_constant__ ChanAttnParams p) {
  constexpr int kRow = kHD + 4;  // 144-byte rows: float4-aligned, and the per-token scatter of the load phase is 4-way
                                 // instead of 32-way bank-conflicted
  __shared__ __align__(16) float qs[kTok][kRow], ks[kTok][kRow];
  __shared__ float red[3][kHD * kHD + 2 * kHD];
  const int h = blockIdx.y, n = blockIdx.z, d = p.head_dim;
  const size_t hw = (size_t)p.H * p.W;
  const size_t chunk = (hw + gridDim.x - 1) / gridDim.x;
  const size_t t0 = (size_t)blockIdx.x * chunk, t1 = min(hw, t0 + chunk);
  const T* src = reinterpret_cast<const T*>(p.src);
  const int cq = p.src_ch_off + h * d, ck = cq + p.qkv_stride;
  const int grp = threadIdx.x >> 6, tile = threadIdx.x & 63;
  const int i0 = (tile >> 3) * 4, j0 = (tile & 7) * 4;
  float acc[4][4] = {};
  float nq[4] = {0, 0, 0, 0}, nk[4] = {0, 0, 0, 0};
  // planes touched by the head's channel range (at most 5 for head_dim <= 32)
  const int pq0 = cq >> 3, npq = ((cq + d - 1) >> 3) - pq0 + 1;
  const int pk0 = ck >> 3, npk = ((ck + d - 1) >> 3) - pk0 + 1;
  for (int e = threadIdx.x; e < kTok * kRow; e += blockDim.x) (&qs[0][0])[e] = 0.0f, (&ks[0][0])[e] = 0.0f;  // padded columns stay zero
  __syncthreads();
  for (size_t base = t0; base < t1; base += kTok) {
    for (int e = threadIdx.x; e < kTok * (npq + npk); e += blockDim.x) {
      const int pl = e / kTok, tt = e - pl * kTok;
      const bool isk = pl >= npq;
      const int plane = isk ? pk0 + (pl - npq) : pq0 + pl;
      const int c0 = isk ? ck : cq;
      const size_t tok = base + tt;
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (tok < t1) load8<T>(src + (((size_t)n * p.src_planes + plane) * hw + tok) * 8, v);
      float* row = isk ? ks[tt] : qs[tt];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = plane * 8 + k - c0;
        if (c >= 0 && c < d) row[c] = v[k];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int tt = grp; tt < kTok; tt += 4) {
      const float4 a = *reinterpret_cast<const float4*>(&qs[tt][i0]);
      const float4 b = *reinterpret_cast<const float4*>(&ks[tt][j0]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        nq[i] = fmaf(av[i], av[i], nq[i]);
        nk[i] = fmaf(bv[i], bv[i], nk[i]);
      }
    }
    __syncthreads();
  }
  // reduce the four token groups in fixed order: groups 1..3 park their tiles, group 0 adds them up and writes
  const int len = d * d + 2 * d;
  if (grp > 0) {
    float* r = red[grp - 1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) r[(i0 + i) * kHD + j0 + j] = acc[i][j];
    if (j0 == 0)
#pragma unroll
      for (int i = 0; i < 4; ++i) r[kHD * kHD + i0 + i] = nq[i];
    if (i0 == 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) r[kHD * kHD + kHD + j0 + j] = nk[j];
  }
  __syncthreads();
  if (grp == 0) {
    float* out = p.partial + (((size_t)n * p.heads + h) * gridDim.x + blockIdx.x) * len;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = (i0 + i) * kHD + j0 + j;
        const float v = ((acc[i][j] + red[0][e]) + red[1][e]) + red[2][e];
        if (i0 + i < d && j0 + j < d) out[(i0 + i) * d + j0 + j] = v;
      }
    if (j0 == 0)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = kHD * kHD + i0 + i;
        if (i0 + i < d) out[d * d + i0 + i] = ((nq[i] + red[0][e]) + red[1][e]) + red[2][e];
      }
    if (i0 == 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = kHD * kHD + kHD + j0 + j;
        if (j0 + j < d) out[d * d + d + j0 + j] = ((nk[j] + red[0][e]) + red[1][e]) + red[2][e];
      }
  }
}

// bf16 plan: the same reduction on warp-level tensor-core MMAs.  G = Q^T K contracts over TOKENS, so both operands are needed
// transposed relative to their natural [token][channel] staging: ldmatrix.trans delivers exactly those fragments.  A CTA stages
// 128 tokens of the head's q and k channels (bf16, 80-byte rows, head dims d..31 zero) and its 8 warps each own one 16 x 8 tile
// of the 32 x 32 Gram matrix — and of Q^T Q and K^T K, whose diagonals are the squared column norms — three MMAs per 16-token
// step.  The CUDA-core kernel above spends ~1 700 thread instructions per token and head (211 us per launch on DAT 4x 512^2,
// 0.15 of the HBM roofline); this one is bound by the staging loads.  Fixed summation order: run-to-run deterministic.
constexpr int kCaTok = 128;          // tokens per staged tile
constexpr int kCaRow = kHD + 8;      // bf16 elements per staged row (80 bytes: 16-byte aligned, conflict-free for ldmatrix)
__global__ void __launch_bounds__(256) chanattn_reduce_mma_kernel(const __grid_constant__ ChanAttnParams p) {
  using T = __nv_bfloat16;
  __shared__ __align__(16) T qs[kCaTok][kCaRow], ks[kCaTok][kCaRow];
  const int h = blockIdx.y, n = blockIdx.z, d = p.head_dim;
  const size_t hw = (size_t)p.H * p.W;
  const size_t chunk = ((hw + gridDim.x - 1) / gridDim.x + kCaTok - 1) / kCaTok * kCaTok;
  const size_t t0 = (size_t)blockIdx.x * chunk, t1 = min(hw, t0 + chunk);
  const T* src = reinterpret_cast<const T*>(p.src);
  const int cq = p.src_ch_off + h * d, ck = cq + p.qkv_stride;
  const int pq0 = cq >> 3, npq = ((cq + d - 1) >> 3) - pq0 + 1;
  const int pk0 = ck >> 3, npk = ((ck + d - 1) >> 3) - pk0 + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp >> 2, nt = warp & 3;  // this warp's tile: rows i in [16 mt, 16 mt + 16), columns j in [8 nt, 8 nt + 8)
  float gqk[4] = {0, 0, 0, 0}, gqq[4] = {0, 0, 0, 0}, gkk[4] = {0, 0, 0, 0};
  for (int e = threadIdx.x; e < kCaTok * kCaRow / 8; e += blockDim.x)
    reinterpret_cast<uint4*>(&qs[0][0])[e] = make_uint4(0, 0, 0, 0), reinterpret_cast<uint4*>(&ks[0][0])[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // ldmatrix row addresses: lanes 0-7 / 8-15 / 16-23 / 24-31 give the rows of the four 8 x 8 matrices
  //   A (x4): (tokens 0-7, i 0-7), (tokens 0-7, i 8-15), (tokens 8-15, i 0-7), (tokens 8-15, i 8-15)  = a0, a1, a2, a3
  //   B (x2): (tokens 0-7, j 0-7), (tokens 8-15, j 0-7)                                              = b0, b1
  const int a_tok = (lane & 7) + ((lane >> 4) << 3), a_col = mt * 16 + (((lane >> 3) & 1) << 3);
  const int b_tok = lane & 15, b_col = nt * 8;
  // Staging: thread = (token tt, plane phase); its items (pl = phase, phase + 2, ...) are the same for every tile.  All loads of
  // a tile are issued back to back, and the NEXT tile's loads are in flight while this tile's MMAs run (one load per loop
  // iteration, waited for before the scatter, made the first version latency-bound: 156 us per launch).
  constexpr int kItems = 5;  // ceil(10 planes / 2 phases): a head of <= 32 channels straddles <= 5 planes per matrix
  const int tt = threadIdx.x & (kCaTok - 1), phase = threadIdx.x >> 7;
  const int nitems = npq + npk;
  uint4 v[kItems];
  auto fetch = [&](size_t base) {
    const size_t tok = base + tt;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int pl = phase + 2 * it;
      v[it] = make_uint4(0, 0, 0, 0);  // tokens beyond the range contribute zeros
      if (pl < nitems && tok < t1) {
        const int plane = pl >= npq ? pk0 + (pl - npq) : pq0 + pl;
        v[it] = *reinterpret_cast<const uint4*>(src + (((size_t)n * p.src_planes + plane) * hw + tok) * 8);
      }
    }
  };
  if (t0 < t1) fetch(t0);
  for (size_t base = t0; base < t1; base += kCaTok) {
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int pl = phase + 2 * it;
      if (pl < nitems) {
        const bool isk = pl >= npq;
        const int plane = isk ? pk0 + (pl - npq) : pq0 + pl;
        const int cb = plane * 8 - (isk ? ck : cq);  // head dim of the plane's first channel
        const T* ev = reinterpret_cast<const T*>(&v[it]);
        T* row = isk ? ks[tt] : qs[tt];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if ((unsigned)(cb + k) < (unsigned)d) row[cb + k] = ev[k];
      }
    }
    __syncthreads();
    if (base + kCaTok < t1) fetch(base + kCaTok);
#pragma unroll
    for (int k16 = 0; k16 < kCaTok / 16; ++k16) {
      uint32_t aq[4], ak[4], bq[2], bk[2];
      ldsm_x4_trans(aq, &qs[k16 * 16 + a_tok][a_col]);
      ldsm_x4_trans(ak, &ks[k16 * 16 + a_tok][a_col]);
      ldsm_x2_trans(bk, &ks[k16 * 16 + b_tok][b_col]);
      ldsm_x2_trans(bq, &qs[k16 * 16 + b_tok][b_col]);
      mma_bf16_16816(gqk, aq, bk[0], bk[1]);
      mma_bf16_16816(gqq, aq, bq[0], bq[1]);
      mma_bf16_16816(gkk, ak, bk[0], bk[1]);
    }
    __syncthreads();
  }
  const int len = d * d + 2 * d;
  float* out = p.partial + (((size_t)n * p.heads + h) * gridDim.x + blockIdx.x) * len;
  const int g4 = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = mt * 16 + g4 + ((e >> 1) << 3), j = nt * 8 + 2 * t4 + (e & 1);
    if (i < d && j < d) {
      out[i * d + j] = gqk[e];
      if (i == j) out[d * d + i] = gqq[e], out[d * d + d + i] = gkk[e];
    }
  }
}

// attn[n][head][i][j] = softmax_j( G[i][j] / (max(|q_i|, eps) max(|k_j|, eps)) * temperature[head] )
__global__ void __launch_bounds__(1024) chanattn_finalize_kernel(const __grid_constant__ ChanAttnParams p) {
  __shared__ float g[kHD * kHD + 2 * kHD];
  const int h = blockIdx.x, n = blockIdx.y, d = p.head_dim;
  const int len = d * d + 2 * d;
  const float* in = p.partial + ((size_t)n * p.heads + h) * p.blocks * len;
  // one element per thread (the 6-CTA, 256-thread version walked four elements x 128 partials per thread: 94 us per launch);
  // fixed order: run-to-run deterministic; 16 loads in flight
  for (int e = threadIdx.x; e < len; e += blockDim.x) {
    double s = 0.0;
    for (int b0 = 0; b0 < p.blocks; b0 += 16) {
      float t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = b0 + u < p.blocks ? in[(size_t)(b0 + u) * len + e] : 0.0f;
#pragma unroll
      for (int u = 0; u < 16; ++u) s += t[u];
    }
    g[e] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x < d) {
    const int i = threadIdx.x;
    const float qn = fmaxf(sqrtf(g[d * d + i]), 1e-12f);
    const float temp = p.temperature[h];
    float row[kHD];
    float mx = -INFINITY;
    for (int j = 0; j < d; ++j) {
      const float kn = fmaxf(sqrtf(g[d * d + d + j]), 1e-12f);
      row[j] = g[i * d + j] / (qn * kn) * temp;
      mx = fmaxf(mx, row[j]);
    }
    float sum = 0.0f;
    for (int j = 0; j < d; ++j) {
      row[j] = expf(row[j] - mx);
      sum += row[j];
    }
    float* out = p.attn + (((size_t)n * p.heads + h) * d + i) * d;
    for (int j = 0; j < d; ++j) out[j] = row[j] / sum;
  }
}

// out[token][head*d + i] = sum_j attn[head][i][j] * v[token][head*d + j]
template <typename T>
__global__ void __launch_bounds__(256) chanattn_apply_kernel(const __grid_constant__ ChanAttnParams p) {
  __shared__ float a[kHD * kHD];
  const int h = blockIdx.y, n = blockIdx.z, d = p.head_dim;
  const size_t hw = (size_t)p.H * p.W;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) a[e] = p.attn[((size_t)n * p.heads + h) * d * d + e];
  __syncthreads();
  const size_t tok = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= hw) return;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  const int cv = p.src_ch_off + 2 * p.qkv_stride + h * d;
  float v[kHD];
#pragma unroll
  for (int c = 0; c < kHD; ++c) {
    const int ch = cv + c;
    v[c] = c < d ? (float)src[(((size_t)n * p.src_planes + (ch >> 3)) * hw + tok) * 8 + (ch & 7)] : 0.0f;
  }
  const int co = p.dst_ch_off + h * d;
  for (int i = 0; i < d; ++i) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < kHD; ++j)
      if (j < d) s = fmaf(a[i * d + j], v[j], s);
    const int ch = co + i;
    dst[(((size_t)n * p.dst_planes + (ch >> 3)) * hw + tok) * 8 + (ch & 7)] = (T)s;
  }
}

// bf16 plan: the per-token 30 x 30 apply on warp-level tensor-core MMAs.  out[tok][i] = sum_j attn[i][j] v[tok][j] is a
// [tokens x d] x [d x d] GEMM per head: a warp takes 16 tokens, its A fragments (tokens x head dims) are 4-byte loads straight
// from the planar buffer (8 neighbouring tokens x 16 bytes per request), the B fragments (the head's softmaxed attention matrix,
// rounded to bf16 like P in the window kernel) live in 16 registers for the CTA's lifetime, results leave as 4-byte bf16 pairs.
// The CUDA-core version above spends one shared-memory load per FMA (251 us per launch on DAT 4x 512^2, 85 % issue-bound).
// Needs even head_dim and even channel offsets (pairs must not straddle a plane); anything else stays on the kernel above.
__global__ void __launch_bounds__(256) chanattn_apply_mma_kernel(const __grid_constant__ ChanAttnParams p) {
  using T = __nv_bfloat16;
  const int h = blockIdx.y, n = blockIdx.z, d = p.head_dim;
  const size_t hw = (size_t)p.H * p.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g4 = lane >> 2, t4 = lane & 3;
  const float* a = p.attn + ((size_t)n * p.heads + h) * d * d;
  uint32_t bf[4][2][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = nt * 8 + g4, j = ks * 16 + 2 * t4 + 8 * half;
        const float lo = (i < d && j < d) ? a[i * d + j] : 0.0f, hi = (i < d && j + 1 < d) ? a[i * d + j + 1] : 0.0f;
        bf[nt][ks][half] = pack_bf16x2(lo, hi);
      }
  const T* src = reinterpret_cast<const T*>(p.src) + (size_t)n * p.src_planes * hw * 8;
  T* dst = reinterpret_cast<T*>(p.dst) + (size_t)n * p.dst_planes * hw * 8;
  const int cv = p.src_ch_off + 2 * p.qkv_stride + h * d, co = p.dst_ch_off + h * d;
  size_t aoff[2][2], ooff[4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ch = cv + ks * 16 + 2 * t4 + 8 * half;
      aoff[ks][half] = (size_t)(ch >> 3) * hw * 8 + (ch & 7);
    }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int ch = co + nt * 8 + 2 * t4;
    ooff[nt] = (size_t)(ch >> 3) * hw * 8 + (ch & 7);
  }
  const size_t tiles = (hw + 15) / 16;
  for (size_t tile = (size_t)blockIdx.x * 8 + warp; tile < tiles; tile += (size_t)gridDim.x * 8) {
    const size_t t0 = tile * 16 + g4, t1 = t0 + 8;
    const bool v0 = t0 < hw, v1 = t1 < hw;
    uint32_t af[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const bool in = ks * 16 + 2 * t4 + 8 * half < d;  // head dims >= d belong to the next head: read them as zero
        af[ks][half * 2 + 0] = (v0 && in) ? *reinterpret_cast<const uint32_t*>(src + aoff[ks][half] + t0 * 8) : 0u;
        af[ks][half * 2 + 1] = (v1 && in) ? *reinterpret_cast<const uint32_t*>(src + aoff[ks][half] + t1 * 8) : 0u;
      }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) mma_bf16_16816(acc[nt], af[ks], bf[nt][ks][0], bf[nt][ks][1]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt * 8 + 2 * t4 >= d) continue;
      if (v0) *reinterpret_cast<uint32_t*>(dst + ooff[nt] + t0 * 8) = pack_bf16x2(acc[nt][0], acc[nt][1]);
      if (v1) *reinterpret_cast<uint32_t*>(dst + ooff[nt] + t1 * 8) = pack_bf16x2(acc[nt][2], acc[nt][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ AIM
// pool partial sums: partial[n][block][plane*8 + k]
template <typename T>
__global__ void __launch_bounds__(256) aim_pool_kernel(const __grid_constant__ AimParams p) {
  __shared__ float red[256][8];
  const int pl = blockIdx.y, n = blockIdx.z;
  const size_t hw = (size_t)p.H * p.W;
  const T* base = reinterpret_cast<const T*>(p.pool_src) + ((size_t)n * p.pool_planes + p.pool_plane0 + pl) * hw * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // four independent 16-byte loads in flight per thread (one at a time: 3.2 TB/s at 19 % issue-active — waiting on HBM latency)
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += 4 * step) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * step < hw) load8<T>(base + (i + u * step) * 8, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * step < hw) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[u][k];
      }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[threadIdx.x][k] += red[threadIdx.x + o][k];
    __syncthreads();
  }
  if (threadIdx.x < 8) p.partial[((size_t)n * gridDim.x + blockIdx.x) * p.cpad + pl * 8 + threadIdx.x] = red[0][threadIdx.x];
}

// cmap[n][c] = sigmoid( W2 . gelu(W1 . mean + b1) + b2 )     (BatchNorm already folded into W1/b1)
__global__ void __launch_bounds__(256) aim_cmap_kernel(const __grid_constant__ AimParams p) {
  __shared__ float mean[512], hid[64];
  const int n = blockIdx.x, C = p.channels;
  const float inv = 1.0f / ((float)p.H * (float)p.W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;  // fixed order, 16 loads in flight
    for (int b0 = 0; b0 < p.blocks; b0 += 16) {
      float t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = b0 + u < p.blocks ? p.partial[((size_t)n * p.blocks + b0 + u) * p.cpad + c] : 0.0f;
#pragma unroll
      for (int u = 0; u < 16; ++u) s += t[u];
    }
    mean[c] = (float)s * inv;
  }
  __syncthreads();
  // hidden layer: one warp per hidden unit, lanes stride over the channels (coalesced weight rows), shuffle tree (fixed order)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < p.ci_hidden; k += blockDim.x >> 5) {
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s = fmaf(p.ci_w1[k * C + c], mean[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hid[k] = gelu_f(s + p.ci_b1[k]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = p.ci_b2[c];
#pragma unroll 8
    for (int k = 0; k < p.ci_hidden; ++k) s = fmaf(p.ci_w2[c * p.ci_hidden + k], hid[k], s);
    p.cmap[(size_t)n * p.cpad + c] = sigm_f(s);
  }
}

// per pixel: smap = sigmoid(w2 . gelu(W1 . s + b1) + b2) on the spatial-map source, then the gated sum
//   mode 0 (window-attention block): Y = ATT * cmap[c] + smap * CONVX      (arch.py:503-508)
//   mode 1 (channel-attention block): Y = ATT * smap + CONVX * cmap[c]     (arch.py:602-607)
template <typename T>
__global__ void __launch_bounds__(256) aim_combine_kernel(const __grid_constant__ AimParams p) {
  extern __shared__ __align__(16) float sm[];
  const int C = p.channels, planes = (C + 7) >> 3, hidn = p.si_hidden;
  float* w1 = sm;                 // [hidn][cpad]
  float* cm = w1 + hidn * p.cpad; // [cpad]
  const int n = blockIdx.y;
  if (p.hid == nullptr)
    for (int e = threadIdx.x; e < hidn * p.cpad; e += blockDim.x) {
      const int k = e / p.cpad, c = e - k * p.cpad;
      w1[e] = c < C ? p.si_w1[k * C + c] : 0.0f;
    }
  for (int c = threadIdx.x; c < p.cpad; c += blockDim.x) cm[c] = c < C ? p.cmap[(size_t)n * p.cpad + c] : 0.0f;
  __syncthreads();
  const size_t hw = (size_t)p.H * p.W;
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= hw) return;
  const T* att = reinterpret_cast<const T*>(p.att) + ((size_t)n * p.att_planes + p.att_plane0) * hw * 8 + pix * 8;
  const T* cvx = reinterpret_cast<const T*>(p.convx) + ((size_t)n * p.convx_planes + p.convx_plane0) * hw * 8 + pix * 8;
  const T* ssrc = p.mode == 0 ? att : cvx;
  float hid[16];
  float s = p.si_b2;
  if (p.hid != nullptr) {
    // the hidden layer gelu(W1 . s + b1) was computed by a 1x1 conv on the tensor cores (1 980 FMAs per pixel here otherwise:
    // the kernel is issue-bound, 51.6 M instructions for 293 MB)
    if constexpr (sizeof(T) == 2) {
      const T* hp = reinterpret_cast<const T*>(p.hid) + ((size_t)n * p.hid_planes * hw + pix) * 8;
      load8<T>(hp, *reinterpret_cast<float(*)[8]>(&hid[0]));
      load8<T>(hp + hw * 8, *reinterpret_cast<float(*)[8]>(&hid[8]));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < hidn) s = fmaf(p.si_w2[k], hid[k], s);
  } else {
#pragma unroll
  for (int k = 0; k < 16; ++k) hid[k] = k < hidn ? p.si_b1[k] : 0.0f;
  // planes in batches of four: four independent 16-byte loads in flight per thread (one at a time left the kernel waiting on
  // HBM latency: 2.4 TB/s)
  for (int pl0 = 0; pl0 < planes; pl0 += 4) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (pl0 + u < planes) load8<T>(ssrc + (size_t)(pl0 + u) * hw * 8, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pl0 + u >= planes) break;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < hidn) {
          // two 16-byte broadcast loads per 8 FMAs (one 4-byte shared-memory load per FMA made this kernel shared-memory bound)
          const float4* wr = reinterpret_cast<const float4*>(w1 + k * p.cpad + (pl0 + u) * 8);
          const float4 wa = wr[0], wb = wr[1];
          hid[k] = fmaf(wa.x, v[u][0], fmaf(wa.y, v[u][1], fmaf(wa.z, v[u][2], fmaf(wa.w, v[u][3], hid[k]))));
          hid[k] = fmaf(wb.x, v[u][4], fmaf(wb.y, v[u][5], fmaf(wb.z, v[u][6], fmaf(wb.w, v[u][7], hid[k]))));
        }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (k < hidn) s = fmaf(p.si_w2[k], gelu_f(hid[k]), s);
  }
  const float smap = sigm_f(s);
  T* dst = reinterpret_cast<T*>(p.dst) + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + pix * 8;
  for (int pl0 = 0; pl0 < planes; pl0 += 2) {
    float a[2][8], b[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (pl0 + u < planes) {
        load8<T>(att + (size_t)(pl0 + u) * hw * 8, a[u]);
        load8<T>(cvx + (size_t)(pl0 + u) * hw * 8, b[u]);
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (pl0 + u >= planes) break;
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float cmv = cm[(pl0 + u) * 8 + c];
        o[c] = p.mode == 0 ? fmaf(a[u][c], cmv, smap * b[u][c]) : fmaf(a[u][c], smap, b[u][c] * cmv);
      }
      store8<T>(dst + (size_t)(pl0 + u) * hw * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ DySample head
// One thread per OUTPUT pixel: for each group, the learned offset of this sub-pixel gives one sampling position in the
// low-res feature map (border-clamped bilinear, exactly grid_sample(align_corners=False, padding_mode='border') after the
// reference's normalise / un-normalise round trip cancels: position = (w + off_x, h + off_y)); the group's channels are
// gathered as 16-byte plane chunks at the four neighbours, blended, and fed straight into the 1x1 end_conv, so the sampled
// high-res feature map (C channels at s^2 times the pixels) is never materialised — only out_ch values per pixel are written.
constexpr int kDyMaxOut = 4, kDyMaxC = 256, kDyMaxOff = 256, kDyMaxGroups = 8;
// offset channel c of the pixel `off` points at; with gate_off > 0 the buffer holds [0.5 * offset | scope] and the gate
// sigmoid(scope) is applied here (one conv op produces both halves)
template <typename T>
__device__ __forceinline__ float dys_offset(const T* off, size_t hw, int c, int gate_off) {
  float v = (float)off[(size_t)(c >> 3) * hw * 8 + (c & 7)];
  if (gate_off > 0) {
    const int cg2 = c + gate_off;
    v *= sigm_f((float)off[(size_t)(cg2 >> 3) * hw * 8 + (cg2 & 7)]);
  }
  return v;
}
template <typename T>
__global__ void __launch_bounds__(256) dysample_kernel(const __grid_constant__ DySampleParams p) {
  __shared__ float w_sm[kDyMaxOut * kDyMaxC];
  __shared__ float ip_sm[kDyMaxOff];
  const int C = p.channels, G = p.groups, s = p.s, s2 = s * s, cg = C / G;
  for (int e = threadIdx.x; e < p.out_ch * C; e += blockDim.x) w_sm[e] = p.weight[e];
  for (int e = threadIdx.x; e < 2 * G * s2; e += blockDim.x) ip_sm[e] = p.init_pos[e];
  __syncthreads();
  const int OH = p.H * s, OW = p.W * s;
  const size_t total = (size_t)p.n * OH * OW;
  const size_t hw = (size_t)p.H * p.W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(idx % OW), Y = (int)((idx / OW) % OH), n = (int)(idx / ((size_t)OW * OH));
    const int w = X / s, j = X - w * s, h = Y / s, i = Y - h * s;
    const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8;
    const T* off = reinterpret_cast<const T*>(p.off) + ((size_t)n * p.off_planes + p.off_plane0) * hw * 8 + ((size_t)h * p.W + w) * 8;
    float acc[kDyMaxOut];
#pragma unroll
    for (int o = 0; o < kDyMaxOut; ++o) acc[o] = o < p.out_ch ? p.bias[o] : 0.0f;
    for (int g = 0; g < G; ++g) {
      const int cx = g * s2 + i * s + j, cy = (G + g) * s2 + i * s + j;
      const float ox = dys_offset<T>(off, hw, cx, p.gate_off) + ip_sm[cx];
      const float oy = dys_offset<T>(off, hw, cy, p.gate_off) + ip_sm[cy];
      const float px = fminf(fmaxf((float)w + ox, 0.0f), (float)(p.W - 1));
      const float py = fminf(fmaxf((float)h + oy, 0.0f), (float)(p.H - 1));
      const int x0 = (int)floorf(px), y0 = (int)floorf(py);
      const int x1 = min(x0 + 1, p.W - 1), y1 = min(y0 + 1, p.H - 1);
      const float fx = px - (float)x0, fy = py - (float)y0;
      const float w00 = (1.0f - fx) * (1.0f - fy), w01 = fx * (1.0f - fy), w10 = (1.0f - fx) * fy, w11 = fx * fy;
      const size_t o00 = ((size_t)y0 * p.W + x0) * 8, o01 = ((size_t)y0 * p.W + x1) * 8, o10 = ((size_t)y1 * p.W + x0) * 8, o11 = ((size_t)y1 * p.W + x1) * 8;
      const int c_lo = g * cg, c_hi = c_lo + cg;
      for (int pl = c_lo >> 3; pl <= (c_hi - 1) >> 3; ++pl) {
        const T* base = src + (size_t)pl * hw * 8;
        float a[8], b[8], c[8], d[8];
        load8<T>(base + o00, a);
        load8<T>(base + o01, b);
        load8<T>(base + o10, c);
        load8<T>(base + o11, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ch = pl * 8 + k;
          if (ch < c_lo || ch >= c_hi) continue;
          const float v = fmaf(w00, a[k], fmaf(w01, b[k], fmaf(w10, c[k], w11 * d[k])));
#pragma unroll
          for (int o = 0; o < kDyMaxOut; ++o)
            if (o < p.out_ch) acc[o] = fmaf(w_sm[o * C + ch], v, acc[o]);
        }
      }
    }
    for (int o = 0; o < p.out_ch; ++o) st_any(p.dst, p.dst_dtype, (((size_t)n * p.out_ch + o) * OH + Y) * OW + X, acc[o]);
  }
}

// Lean version for groups of <= 32 channels (every released model: 12 or 9): ncu showed the kernel above issue-bound (83 % of
// the issue slots, 2 200 instructions per output pixel) on 64-bit index divisions and per-channel group-membership predicates.
// Here the grid is (x blocks, output row, image) — no divisions by the image width — and the end_conv weights are pre-masked
// per (group, plane of the group's span) in shared memory, zero outside the group, so the inner loop is branch-free:
// 4 chunk loads, 8 blends, 8 x 4 FMAs per plane.
constexpr int kDySpan = 5;  // planes a group of <= 32 channels can straddle
template <typename T>
__global__ void __launch_bounds__(256) dysample_lean_kernel(const __grid_constant__ DySampleParams p) {
  __shared__ __align__(16) float wm[kDyMaxGroups * kDySpan * 8 * 4];
  __shared__ float ip_sm[kDyMaxOff];
  const int C = p.channels, G = p.groups, s = p.s, s2 = s * s, cg = C / G;
  for (int e = threadIdx.x; e < G * kDySpan * 8 * 4; e += blockDim.x) {
    const int o = e & 3, k = (e >> 2) & 7, q = (e >> 5) % kDySpan, g = (e >> 5) / kDySpan;
    const int ch = (((g * cg) >> 3) + q) * 8 + k;
    wm[e] = (ch >= g * cg && ch < (g + 1) * cg && o < p.out_ch) ? p.weight[o * C + ch] : 0.0f;
  }
  for (int e = threadIdx.x; e < 2 * G * s2; e += blockDim.x) ip_sm[e] = p.init_pos[e];
  __syncthreads();
  const int OW = p.W * s, OH = p.H * s;
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, n = blockIdx.z;
  if (X >= OW) return;
  const int h = Y / s, i = Y - h * s, w = X / s, j = X - w * s;
  const size_t hw = (size_t)p.H * p.W;
  const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8;
  const T* off = reinterpret_cast<const T*>(p.off) + ((size_t)n * p.off_planes + p.off_plane0) * hw * 8 + ((size_t)h * p.W + w) * 8;
  float acc[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = o < p.out_ch ? p.bias[o] : 0.0f;
  for (int g = 0; g < G; ++g) {
    const int cx = g * s2 + i * s + j, cy = (G + g) * s2 + i * s + j;
    const float ox = dys_offset<T>(off, hw, cx, p.gate_off) + ip_sm[cx];
    const float oy = dys_offset<T>(off, hw, cy, p.gate_off) + ip_sm[cy];
    const float px = fminf(fmaxf((float)w + ox, 0.0f), (float)(p.W - 1));
    const float py = fminf(fmaxf((float)h + oy, 0.0f), (float)(p.H - 1));
    const int x0 = (int)px, y0 = (int)py;  // px, py >= 0: truncation == floor
    const int x1 = min(x0 + 1, p.W - 1), y1 = min(y0 + 1, p.H - 1);
    const float fx = px - (float)x0, fy = py - (float)y0;
    const float w00 = (1.0f - fx) * (1.0f - fy), w01 = fx * (1.0f - fy), w10 = (1.0f - fx) * fy, w11 = fx * fy;
    const size_t o00 = ((size_t)y0 * p.W + x0) * 8, o01 = ((size_t)y0 * p.W + x1) * 8, o10 = ((size_t)y1 * p.W + x0) * 8, o11 = ((size_t)y1 * p.W + x1) * 8;
    const int pl0 = (g * cg) >> 3, span = (((g + 1) * cg - 1) >> 3) - pl0 + 1;
    for (int q = 0; q < span; ++q) {
      const T* bp = src + (size_t)(pl0 + q) * hw * 8;
      float a[8], b[8], c[8], d[8];
      load8<T>(bp + o00, a);
      load8<T>(bp + o01, b);
      load8<T>(bp + o10, c);
      load8<T>(bp + o11, d);
      const float4* wq = reinterpret_cast<const float4*>(wm) + (g * kDySpan + q) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = fmaf(w00, a[k], fmaf(w01, b[k], fmaf(w10, c[k], w11 * d[k])));
        const float4 ww = wq[k];
        acc[0] = fmaf(ww.x, v, acc[0]);
        acc[1] = fmaf(ww.y, v, acc[1]);
        acc[2] = fmaf(ww.z, v, acc[2]);
        acc[3] = fmaf(ww.w, v, acc[3]);
      }
    }
  }
  for (int o = 0; o < p.out_ch; ++o) st_any(p.dst, p.dst_dtype, (((size_t)n * p.out_ch + o) * OH + Y) * OW + X, acc[o]);
}

// Pre-projected form (DySampleParams::projected): sampling and the 1x1 end_conv are both linear, so the end_conv is applied
// FIRST, per group, on the low-res grid by an ordinary tensor-core 1x1 conv op: z[g*4 + o] = sum_{c in group g} W[o][c] x[c]
// (4 channels per group, out_ch of them used).  The head then gathers 4 values per (group, neighbour) — one 8-byte load in
// bf16 — instead of the group's 12-16 feature channels: a quarter of the loads, conversions and FMAs of the kernels above.
// out[o](Y, X) = bias[o] + sum_g bilinear_g(z[g*4 + o]).
template <typename T>
__global__ void __launch_bounds__(256) dysample_proj_kernel(const __grid_constant__ DySampleParams p) {
  __shared__ float ip_sm[kDyMaxOff];
  const int G = p.groups, s = p.s, s2 = s * s;
  for (int e = threadIdx.x; e < 2 * G * s2; e += blockDim.x) ip_sm[e] = p.init_pos[e];
  __syncthreads();
  const int OW = p.W * s, OH = p.H * s;
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y, n = blockIdx.z;
  if (X >= OW) return;
  const int h = Y / s, i = Y - h * s, w = X / s, j = X - w * s;
  const size_t hw = (size_t)p.H * p.W;
  const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8;
  const T* off = reinterpret_cast<const T*>(p.off) + ((size_t)n * p.off_planes + p.off_plane0) * hw * 8 + ((size_t)h * p.W + w) * 8;
  float acc[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = o < p.out_ch ? p.bias[o] : 0.0f;
  for (int g = 0; g < G; ++g) {
    const int cx = g * s2 + i * s + j, cy = (G + g) * s2 + i * s + j;
    const float ox = dys_offset<T>(off, hw, cx, p.gate_off) + ip_sm[cx];
    const float oy = dys_offset<T>(off, hw, cy, p.gate_off) + ip_sm[cy];
    const float px = fminf(fmaxf((float)w + ox, 0.0f), (float)(p.W - 1));
    const float py = fminf(fmaxf((float)h + oy, 0.0f), (float)(p.H - 1));
    const int x0 = (int)px, y0 = (int)py;  // px, py >= 0: truncation == floor
    const int x1 = min(x0 + 1, p.W - 1), y1 = min(y0 + 1, p.H - 1);
    const float fx = px - (float)x0, fy = py - (float)y0;
    const float wt[4] = {(1.0f - fx) * (1.0f - fy), fx * (1.0f - fy), (1.0f - fx) * fy, fx * fy};
    const size_t nb[4] = {((size_t)y0 * p.W + x0) * 8, ((size_t)y0 * p.W + x1) * 8, ((size_t)y1 * p.W + x0) * 8, ((size_t)y1 * p.W + x1) * 8};
    const T* bp = src + (size_t)(g >> 1) * hw * 8 + (g & 1) * 4;  // two groups of 4 channels per 8-channel plane
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float v[4];
      if constexpr (sizeof(T) == 2) {
        const uint2 r = *reinterpret_cast<const uint2*>(bp + nb[t]);
        v[0] = __uint_as_float(r.x << 16), v[1] = __uint_as_float(r.x & 0xFFFF0000u);
        v[2] = __uint_as_float(r.y << 16), v[3] = __uint_as_float(r.y & 0xFFFF0000u);
      } else {
        const float4 r = *reinterpret_cast<const float4*>(bp + nb[t]);
        v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = fmaf(wt[t], v[o], acc[o]);
    }
  }
  for (int o = 0; o < p.out_ch; ++o) st_any(p.dst, p.dst_dtype, (((size_t)n * p.out_ch + o) * OH + Y) * OW + X, acc[o]);
}

inline int grid_for(size_t total, int threads = 256, int cap = 148 * 32) {
  return (int)std::max<size_t>(1, std::min<size_t>((total + threads - 1) / threads, (size_t)cap));
}

}  // namespace

cudaError_t launch_layernorm(const TokenOpParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int sms = num_sms > 0 ? num_sms : 148;
  const int g = grid_for((size_t)p.n * p.H * p.W, 256, sms * 32);
  const int planes = (p.channels + 7) / 8;
  const size_t pixels = (size_t)p.n * p.H * p.W;
  const int g2 = (int)std::min<size_t>((pixels + 63) / 64, (size_t)sms * 128);
  static const bool no_stream = rsb_env("RSB_LN_REG") != nullptr;  // bring-up: the register-resident kernel instead
  if (bf16 && !no_stream && planes <= 64) {
    const size_t stage = (size_t)planes * kLnTile * 16;
    // two resident CTAs per SM with a two-stage ring each when that fits: the kernel is bound by instruction issue / latency, and 18
    // warps hide more of it than 9 (A/B on one box, DAT 4x 512^2: 29.2 -> 24.8 us per LayerNorm launch)
    const int ctas = 2 * stage + 1024 <= 100 * 1024 ? 2 : 1;
    int stages = (int)std::min<size_t>(8, ((ctas == 2 ? 104 : 200) * 1024 - (size_t)planes * 64 - 256) / stage);
    const size_t hw = (size_t)p.H * p.W;
    const size_t tiles = (size_t)p.n * ((hw + kLnTile - 1) / kLnTile);
    if (stages >= 2 && tiles >= 1) {
      const size_t smem = (size_t)stages * stage + (size_t)planes * 64 + (size_t)stages * 16 + 16;
      const int grid = (int)std::min<size_t>(tiles, (size_t)sms * ctas);  // persistent
      layernorm_stream_kernel<<<grid, kLnConsumers + 32, smem, s>>>(p, stages);
      return cudaGetLastError();
    }
  }
  if (bf16 && planes <= 8)
    layernorm_bf16_kernel<2><<<g2, 256, 0, s>>>(p);
  else if (bf16 && planes <= 16)
    layernorm_bf16_kernel<4><<<g2, 256, 0, s>>>(p);
  else if (bf16 && planes <= 24)
    layernorm_bf16_kernel<6><<<g2, 256, 0, s>>>(p);
  else if (bf16 && planes <= 32)
    layernorm_bf16_kernel<8><<<g2, 256, 0, s>>>(p);
  else if (bf16)
    layernorm_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    layernorm_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_dwconv3(const TokenOpParams& p, bool bf16, cudaStream_t s) {
  const int g = grid_for((size_t)p.n * ((p.channels + 7) / 8) * p.H * p.W);
  // (a shared-memory ring version — eight rows x 130 pixels filled by cp.async, four rows in flight, 32 rows per CTA, the 3 x 3
  // neighbourhood from shared memory — measured 76 / 115 us against 67 / 82 us for this register ring on DAT's two depthwise convs)
  if (bf16 && (long long)p.n * ((p.channels + 7) / 8) <= 65535)
    dwconv3_bf16_kernel<<<dim3((p.W + 127) / 128, (p.H + kDwRows - 1) / kDwRows, p.n * ((p.channels + 7) / 8)), 128, 0, s>>>(p);
  else if (bf16)
    dwconv3_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    dwconv3_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

size_t winattn_smem_bytes(int split_h, int split_w) {
  const int N = split_h * split_w;
  return (size_t)N * kHD * 4 * 2 + (size_t)(2 * split_h - 1) * (2 * split_w - 1) * 4 + (size_t)N * 4;
}

cudaError_t winattn_configure() {
  {
    cudaError_t e0 = cudaFuncSetAttribute(layernorm_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e0 != cudaSuccess) return e0;
  }
  cudaError_t e = cudaFuncSetAttribute(winattn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(winattn_mma_kernel<256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(winattn_mma_kernel<128, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);


  if (e != cudaSuccess) return e;
  e = winattn_tc_configure();
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(winattn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

cudaError_t launch_winattn(const WinAttnParams& p, bool bf16, int num_sms, cudaStream_t s) {
  if (p.head_pad != 0) return bf16 ? launch_winattn_tc(p, num_sms, s) : cudaErrorInvalidValue;  // validated in rsb_plan_add_op
  const int windows = (p.Hp / p.split_h) * (p.Wp / p.split_w);
  const dim3 grid(windows * p.n, p.heads / 2, 2);
  const size_t smem = winattn_smem_bytes(p.split_h, p.split_w);
  static const bool no_mma = rsb_env("RSB_WINATTN_SIMT") != nullptr;
  const int N = p.split_h * p.split_w;
  if (bf16 && !no_mma && N <= 256 && p.head_dim <= kHD && winattn_mma_smem_bytes(p.split_h, p.split_w) <= 100 * 1024) {
    const int warps = (N + 15) / 16;
    // > 64 tokens: 256-thread CTAs, every warp walks NQ / 128 query tiles; two CTAs per SM overlap one window's staging with
    // the other's attention.  DAT's 8x32 windows at 4x 512^2: one 512-thread CTA per SM 565 us, two 256-thread CTAs 473 us per
    // launch (three 128-thread CTAs measured the same as two of 256).
    if (warps <= 4)  // (a 6-CTA / 80-register variant spills and measured no faster: 260 vs 254 us)
      winattn_mma_kernel<128, 5, false><<<grid, 128, winattn_mma_smem_bytes(p.split_h, p.split_w), s>>>(p);
    else
      winattn_mma_kernel<256, 2, true><<<grid, 32 * std::min(warps, 8), winattn_mma_smem_bytes(p.split_h, p.split_w), s>>>(p);
  } else if (bf16)
    winattn_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(p);
  else
    winattn_kernel<float><<<grid, 256, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_chanattn(const ChanAttnParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const size_t hw = (size_t)p.H * p.W;
  const dim3 g1(p.blocks, p.heads, p.n), g2(p.heads, p.n), g3((unsigned)((hw + 255) / 256), p.heads, p.n);
  if (bf16 && p.head_dim <= kHD) {
    // tensor-core reduction: two resident CTAs per SM stream disjoint token ranges; fewer partial blocks for the finalize step
    ChanAttnParams q = p;
    q.blocks = std::max(1, std::min(p.blocks, 2 * (num_sms > 0 ? num_sms : 148) / std::max(1, p.heads * p.n)));
    chanattn_reduce_mma_kernel<<<dim3(q.blocks, p.heads, p.n), 256, 0, s>>>(q);
    chanattn_finalize_kernel<<<g2, 1024, 0, s>>>(q);
  } else {
    if (bf16)
      chanattn_reduce_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>(p);
    else
      chanattn_reduce_kernel<float><<<g1, 256, 0, s>>>(p);
    chanattn_finalize_kernel<<<g2, 1024, 0, s>>>(p);
  }
  const bool pairs_ok = p.head_dim % 2 == 0 && p.head_dim <= kHD && (p.src_ch_off + 2 * p.qkv_stride) % 2 == 0 && p.dst_ch_off % 2 == 0;
  if (bf16 && pairs_ok) {
    const unsigned gx = (unsigned)std::min<size_t>((hw + 127) / 128, 296);  // 8 warps x 16 tokens per block step
    chanattn_apply_mma_kernel<<<dim3(gx, p.heads, p.n), 256, 0, s>>>(p);
  } else if (bf16)
    chanattn_apply_kernel<__nv_bfloat16><<<g3, 256, 0, s>>>(p);
  else
    chanattn_apply_kernel<float><<<g3, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_aim(const AimParams& p0, bool bf16, int num_sms, cudaStream_t s) {
  // a few pooling CTAs per SM are enough to stream the map; fewer partial blocks make the single-CTA channel-map step (which adds
  // them up) three times shorter
  AimParams p = p0;
  const int planes = (p.channels + 7) / 8;
  p.blocks = std::max(1, std::min(p0.blocks, std::max(8, 4 * (num_sms > 0 ? num_sms : 148) / std::max(1, planes * p.n))));
  const size_t hw = (size_t)p.H * p.W;
  const dim3 g1(p.blocks, planes, p.n), g3((unsigned)((hw + 255) / 256), p.n);
  const size_t smem = ((size_t)p.si_hidden * p.cpad + p.cpad) * sizeof(float);
  if (bf16)
    aim_pool_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>(p);
  else
    aim_pool_kernel<float><<<g1, 256, 0, s>>>(p);
  aim_cmap_kernel<<<p.n, 256, 0, s>>>(p);
  if (bf16)
    aim_combine_kernel<__nv_bfloat16><<<g3, 256, smem, s>>>(p);
  else
    aim_combine_kernel<float><<<g3, 256, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_dysample(const DySampleParams& p, bool bf16, cudaStream_t s) {
  const int g = grid_for((size_t)p.n * p.H * p.s * p.W * p.s, 256, 148 * 64);
  static const bool generic = rsb_env("RSB_DYS_GENERIC") != nullptr;  // bring-up: the straightforward kernel
  const bool lean = !generic && p.groups <= kDyMaxGroups && p.channels / p.groups <= 32 && p.H * p.s <= 65535 && p.n <= 65535;
  const dim3 gl((unsigned)((p.W * p.s + 255) / 256), (unsigned)(p.H * p.s), (unsigned)p.n);
  if (p.projected) {
    if (bf16)
      dysample_proj_kernel<__nv_bfloat16><<<gl, 256, 0, s>>>(p);
    else
      dysample_proj_kernel<float><<<gl, 256, 0, s>>>(p);
  } else if (bf16 && lean)
    dysample_lean_kernel<__nv_bfloat16><<<gl, 256, 0, s>>>(p);
  else if (lean)
    dysample_lean_kernel<float><<<gl, 256, 0, s>>>(p);
  else if (bf16)
    dysample_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    dysample_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace rsb
