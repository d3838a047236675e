// Token-wise ops of the gated-CNN SPAN descendants (RTMoSR) on planar-8 activations.
//
//   rmsnorm_kernel          channels-first RMSNorm: x / (|x|_2 / sqrt(C) + eps) * scale + offset
//                           (/root/reference/resselt/archs/rtmosr/arch.py:25-38)
//   unshuffle_pool_kernel   ParPixelUnshuffle's two inputs in one pass over the map: PixelUnshuffle(2) (C -> 4C channels on
//                           the half grid, channel c*4 + i*2 + j) and MaxPool2d(2) (C channels on the half grid) (arch.py:284-292)
//   dwconv_k_kernel         depthwise K x K conv (OmniShift's re-parameterised 5 x 5, arch.py:215-281)
//   se_*                    CSELayer (global mean -> 1x1 -> ReLU -> 1x1 -> Hardsigmoid -> channel scale, arch.py:7-21) fused with
//                           the PixelShuffle(2) that follows it in GatedCNNBlock.conv (arch.py:314-319): deterministic two-stage
//                           mean, tiny MLP, then scale + shuffle in one pass
// All arithmetic is fp32 regardless of the storage type (bf16 or fp32): one set of kernels serves both plans.  These are
// bandwidth-bound element-wise / stencil passes; they are written for correctness and coalescing (16-byte plane chunks per thread),
// not yet tuned like the DAT token kernels.
#include <algorithm>

#include "kernels.cuh"

namespace rsb {
namespace {

inline int grid_for(size_t total, int threads, int cap) {
  return (int)std::max<size_t>(1, std::min<size_t>((total + threads - 1) / threads, (size_t)cap));
}

// ------------------------------------------------------------------------------------------------ RMSNorm
template <typename T>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __grid_constant__ TokenOpParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.channels, planes = (C + 7) >> 3;
  const size_t total = (size_t)p.n * hw;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  const float inv_sqrt_c = rsqrtf((float)C);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, pix = i - n * hw;
    const T* s = src + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8 + pix * 8;
    float ss = 0.0f;
    for (int pl = 0; pl < planes; ++pl) {
      float v[8];
      load8<T>(s + (size_t)pl * hw * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pl * 8 + k < C) ss = fmaf(v[k], v[k], ss);
    }
    // x.norm(2, dim=1) * C^-0.5, then x / (rms + eps)
    const float inv = 1.0f / (sqrtf(ss) * inv_sqrt_c + p.f0);
    T* d = dst + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + pix * 8;
    for (int pl = 0; pl < planes; ++pl) {
      float v[8], o[8];
      load8<T>(s + (size_t)pl * hw * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = pl * 8 + k;
        o[k] = c < C ? fmaf(p.w0[c], v[k] * inv, p.w1[c]) : 0.0f;
      }
      store8<T>(d + (size_t)pl * hw * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ PixelUnshuffle(2) + MaxPool2d(2)
// src: C channels on the (2H x 2W) grid; dst: [4C unshuffled | C pooled] on the (H x W) grid.  C % 8 == 0.
// One thread per (half-grid pixel, source plane): four 16-byte loads, four + one 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(256) unshuffle_pool_kernel(const __grid_constant__ TokenOpParams p) {
  const int H = p.H, W = p.W;  // the half (destination) grid
  const size_t hw = (size_t)H * W, hw2 = hw * 4;
  const int planes = p.channels >> 3;
  const size_t total = (size_t)p.n * planes * hw;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int pl = (int)((i / hw) % planes);
    const int n = (int)(i / (hw * planes));
    const T* s = src + ((size_t)n * p.src_planes + p.src_plane0 + pl) * hw2 * 8;
    float q[4][8];  // q[i*2 + j] = the source pixel (2y + i, 2x + j)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) load8<T>(s + ((size_t)(2 * y + a) * (2 * W) + 2 * x + b) * 8, q[a * 2 + b]);
    T* d = dst + ((size_t)n * p.dst_planes + p.dst_plane0) * hw * 8 + ((size_t)y * W + x) * 8;
    // unshuffled channel (pl*8 + cc)*4 + phase lives in destination plane pl*4 + cc/2 at position (cc % 2)*4 + phase
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = q[e & 3][2 * k + (e >> 2)];
      store8<T>(d + (size_t)(pl * 4 + k) * hw * 8, o);
    }
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fmaxf(fmaxf(q[0][e], q[1][e]), fmaxf(q[2][e], q[3][e]));
    store8<T>(d + (size_t)(4 * planes + pl) * hw * 8, m);
  }
}

// ------------------------------------------------------------------------------------------------ depthwise K x K
// One CTA works on one 8-channel plane: its K*K x 8 weights sit in shared memory as [tap][8] (two 16-byte broadcast loads per tap
// instead of eight scalar global loads), one thread per output pixel, the K*K neighbour chunks come through L1.
// (First version read the weights from global memory per FMA group: 839 us for RTMoSR's 128-channel half-resolution map at 1080p.)
template <typename T>
__global__ void __launch_bounds__(256) dwconv_k_kernel(const __grid_constant__ TokenOpParams p, int K) {
  __shared__ __align__(16) float wsm[121 * 8 + 8];
  __shared__ uint32_t taps[121];  // (t | ky << 8 | kx << 16) of the taps with a non-zero weight on any of the plane's channels (GateRV3's inception conv is an 11 x 11
  __shared__ int ntaps;          // kernel with 1 / 9 / 11 live taps per channel: arch.py:527-557)
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.channels, R = K / 2, KK = K * K;
  const int pl = blockIdx.y, n = blockIdx.z;
  for (int e = threadIdx.x; e < KK * 8 + 8; e += blockDim.x) {
    const int t = e >> 3, k = e & 7, c = pl * 8 + k;
    wsm[e] = c < C ? (t < KK ? p.w0[c * KK + t] : p.w1[c]) : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0;
    for (int t = 0; t < KK; ++t) {
      bool live = false;
      for (int k = 0; k < 8; ++k) live |= wsm[t * 8 + k] != 0.0f;
      if (live) taps[m++] = (uint32_t)t | (uint32_t)(t / K) << 8 | (uint32_t)(t % K) << 16;
    }
    ntaps = m;
  }
  __syncthreads();
  const int nt = ntaps;
  const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0 + pl) * hw * 8;
  T* dst = reinterpret_cast<T*>(p.dst) + ((size_t)n * p.dst_planes + p.dst_plane0 + pl) * hw * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / p.W), x = (int)(i - (size_t)y * p.W);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = wsm[KK * 8 + k];
    if (nt == KK) {  // dense kernel (RTMoSR's OmniShift 5 x 5): plain loops, no tap list
      for (int ky = 0; ky < K; ++ky) {
        const int sy = y + ky - R;
        if (sy < 0 || sy >= p.H) continue;
        for (int kx = 0; kx < K; ++kx) {
          const int sx = x + kx - R;
          if (sx < 0 || sx >= p.W) continue;
          float v[8];
          load8<T>(src + ((size_t)sy * p.W + sx) * 8, v);
          const float4 wa = *reinterpret_cast<const float4*>(&wsm[(ky * K + kx) * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&wsm[(ky * K + kx) * 8 + 4]);
          acc[0] = fmaf(v[0], wa.x, acc[0]), acc[1] = fmaf(v[1], wa.y, acc[1]), acc[2] = fmaf(v[2], wa.z, acc[2]), acc[3] = fmaf(v[3], wa.w, acc[3]);
          acc[4] = fmaf(v[4], wb.x, acc[4]), acc[5] = fmaf(v[5], wb.y, acc[5]), acc[6] = fmaf(v[6], wb.z, acc[6]), acc[7] = fmaf(v[7], wb.w, acc[7]);
        }
      }
    } else
    for (int ti = 0; ti < nt; ++ti) {
      const uint32_t tp = taps[ti];
      const int t = tp & 0xFF, ky = (tp >> 8) & 0xFF, kx = tp >> 16;
      const int sy = y + ky - R, sx = x + kx - R;
      if (sy < 0 || sy >= p.H || sx < 0 || sx >= p.W) continue;
      float v[8];
      load8<T>(src + ((size_t)sy * p.W + sx) * 8, v);
      const float4 wa = *reinterpret_cast<const float4*>(&wsm[t * 8]);
      const float4 wb = *reinterpret_cast<const float4*>(&wsm[t * 8 + 4]);
      acc[0] = fmaf(v[0], wa.x, acc[0]), acc[1] = fmaf(v[1], wa.y, acc[1]), acc[2] = fmaf(v[2], wa.z, acc[2]), acc[3] = fmaf(v[3], wa.w, acc[3]);
      acc[4] = fmaf(v[4], wb.x, acc[4]), acc[5] = fmaf(v[5], wb.y, acc[5]), acc[6] = fmaf(v[6], wb.z, acc[6]), acc[7] = fmaf(v[7], wb.w, acc[7]);
    }
    store8<T>(dst + i * 8, acc);
  }
}

// Dense depthwise 5 x 5 (RTMoSR's re-parameterised OmniShift, 128 channels on the half grid: 25 % of the model's time in the generic
// kernel above, which loads every input pixel 25 times and spends ~500 instructions per output).  Scatter form: a thread walks down
// one pixel column; every input row (five 16-byte loads, unpacked once) is accumulated into the FIVE output rows it contributes to,
// kept as a ring of accumulators whose slots are compile-time (the row loop is unrolled by five); the row whose last contribution
// has arrived is stored.  5 loads, 100 packed FMAs and 50 shared-memory weight loads per output instead of 25 / 200 / 50.
constexpr int kDw5Rows = 31;  // output rows per CTA: 31 + 4 input rows = 7 rounds of 5
__device__ __forceinline__ void fma2_rt(float& a0, float& a1, float v0, float v1, float w0, float w1) {
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
template <typename T>
__global__ void __launch_bounds__(128) dwconv5_rows_kernel(const __grid_constant__ TokenOpParams p) {
  __shared__ __align__(16) float wsm[25 * 8 + 8];
  const size_t hw = (size_t)p.H * p.W;
  const int C = p.channels;
  const int pl = blockIdx.z % ((C + 7) >> 3), n = blockIdx.z / ((C + 7) >> 3);
  for (int e = threadIdx.x; e < 25 * 8 + 8; e += blockDim.x) {
    const int t = e >> 3, k = e & 7, c = pl * 8 + k;
    wsm[e] = c < C ? (t < 25 ? p.w0[c * 25 + t] : p.w1[c]) : 0.0f;
  }
  __syncthreads();
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= p.W) return;
  const int yb = blockIdx.y * kDw5Rows;
  const T* src = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0 + pl) * hw * 8;
  T* dst = reinterpret_cast<T*>(p.dst) + ((size_t)n * p.dst_planes + p.dst_plane0 + pl) * hw * 8;
  float bias[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bias[k] = wsm[25 * 8 + k];
  float acc[5][8];
#pragma unroll
  for (int s5 = 0; s5 < 5; ++s5)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[s5][k] = bias[k];
  // input row sy = yb - 2 + 5 * round + u contributes through kernel row ky to output row sy + 2 - ky, ring slot (u + 2 - ky) mod 5
  for (int round = 0; round < (kDw5Rows + 4) / 5; ++round) {
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int sy = yb - 2 + 5 * round + u;
      if (sy >= 0 && sy < p.H) {
        const T* row = src + (size_t)sy * p.W * 8;
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const int sx = x + kx - 2;
          if (sx < 0 || sx >= p.W) continue;
          float v[8];
          load8<T>(row + (size_t)sx * 8, v);
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            const float4 wa = *reinterpret_cast<const float4*>(&wsm[(ky * 5 + kx) * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&wsm[(ky * 5 + kx) * 8 + 4]);
            float(&a)[8] = acc[(u + 2 - ky + 5) % 5];
            fma2_rt(a[0], a[1], v[0], v[1], wa.x, wa.y);
            fma2_rt(a[2], a[3], v[2], v[3], wa.z, wa.w);
            fma2_rt(a[4], a[5], v[4], v[5], wb.x, wb.y);
            fma2_rt(a[6], a[7], v[6], v[7], wb.z, wb.w);
          }
        }
      }
      // output row sy - 2 has received its last contribution (ky = 4)
      const int oy = sy - 2;
      float(&done)[8] = acc[(u + 3) % 5];
      if (oy >= yb && oy < yb + kDw5Rows && oy < p.H) store8<T>(dst + ((size_t)oy * p.W + x) * 8, done);
#pragma unroll
      for (int k = 0; k < 8; ++k) done[k] = bias[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ SE + PixelShuffle(2)
// partial[n][block][c]: per-channel sums over this block's pixel range (fixed order: deterministic)
template <typename T>
__global__ void __launch_bounds__(256) se_pool_kernel(const __grid_constant__ SeParams p) {
  __shared__ float red[256][8];
  const int pl = blockIdx.y, n = blockIdx.z;
  const size_t hw = (size_t)p.H * p.W;
  const T* base = reinterpret_cast<const T*>(p.src) + ((size_t)n * p.src_planes + p.src_plane0 + pl) * hw * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    load8<T>(base + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
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
  if (threadIdx.x < 8) p.partial[((size_t)n * gridDim.x + blockIdx.x) * p.channels + pl * 8 + threadIdx.x] = red[0][threadIdx.x];
}

// gate[n][c] = hardsigmoid(W2 . relu(W1 . mean + b1) + b2)
__global__ void __launch_bounds__(256) se_gate_kernel(const __grid_constant__ SeParams p) {
  extern __shared__ float sm[];
  float* mean = sm;          // [C]
  float* hid = sm + p.channels;  // [hidden]
  const int n = blockIdx.x, C = p.channels;
  const float inv = 1.0f / ((float)p.H * (float)p.W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < p.blocks; ++b) s += p.partial[((size_t)n * p.blocks + b) * C + c];
    mean[c] = (float)s * inv;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < p.hidden; k += blockDim.x) {
    float s = p.b1[k];
    for (int c = 0; c < C; ++c) s = fmaf(p.w1[k * C + c], mean[c], s);
    hid[k] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = p.b2[c];
    for (int k = 0; k < p.hidden; ++k) s = fmaf(p.w2[c * p.hidden + k], hid[k], s);
    p.gate[(size_t)n * C + c] = fminf(fmaxf(s + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);  // nn.Hardsigmoid
  }
}

// dst (C/4 channels on the 2H x 2W grid): channel c at (2y + i, 2x + j) = src channel c*4 + i*2 + j at (y, x) [* gate]
template <typename T>
__global__ void __launch_bounds__(256) se_shuffle_kernel(const __grid_constant__ SeParams p) {
  const int H = p.H, W = p.W;  // the source (half) grid
  const size_t hw = (size_t)H * W, hw2 = hw * 4;
  const int oplanes = p.channels >> 5;  // destination planes: C / 4 channels / 8
  const size_t total = (size_t)p.n * oplanes * hw;
  const T* src = reinterpret_cast<const T*>(p.src);
  T* dst = reinterpret_cast<T*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int pl = (int)((i / hw) % oplanes);
    const int n = (int)(i / (hw * oplanes));
    const T* s = src + ((size_t)n * p.src_planes + p.src_plane0) * hw * 8 + ((size_t)y * W + x) * 8;
    float q[4][8];  // q[phase][cc]: destination channel pl*8 + cc, phase i*2 + j
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      load8<T>(s + (size_t)(pl * 4 + k) * hw * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int sc = (pl * 4 + k) * 8 + e;  // source channel = (pl*8 + cc)*4 + phase with cc = 2k + e/4, phase = e % 4
        q[e & 3][2 * k + (e >> 2)] = p.gate != nullptr ? v[e] * p.gate[(size_t)n * p.channels + sc] : v[e];
      }
    }
    T* d = dst + ((size_t)n * p.dst_planes + p.dst_plane0 + pl) * hw2 * 8;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) store8<T>(d + ((size_t)(2 * y + a) * (2 * W) + 2 * x + b) * 8, q[a * 2 + b]);
  }
}

// ------------------------------------------------------------------------------------------------ simplified channel attention
// GateRV3 MetaGated (/root/reference/resselt/archs/gaterv3/arch.py:640-667): x * sca(x) * gamma0 + short with
// sca = Conv1x1(AdaptiveAvgPool2d(1)(x)).  Pooling re-uses se_pool_kernel; gate[n][c] = (W . mean + b)[c] * gamma0[c].
__global__ void __launch_bounds__(256) sca_gate_kernel(const __grid_constant__ SeParams p) {
  extern __shared__ float sm[];
  float* mean = sm;  // [C]
  const int n = blockIdx.x, C = p.channels;
  const float inv = 1.0f / ((float)p.H * (float)p.W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < p.blocks; ++b) s += p.partial[((size_t)n * p.blocks + b) * C + c];
    mean[c] = (float)s * inv;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = p.b1[c];
    for (int k = 0; k < C; ++k) s = fmaf(p.w1[c * C + k], mean[k], s);
    p.gate[(size_t)n * C + c] = s * p.w2[c];
  }
}

// dst = src * gate[n][c] + res
template <typename T>
__global__ void __launch_bounds__(256) sca_apply_kernel(const __grid_constant__ SeParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const int planes = p.channels >> 3;
  const size_t total = (size_t)p.n * planes * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t px = i % hw;
    const int pl = (int)((i / hw) % planes), n = (int)(i / (hw * planes));
    float v[8], r[8];
    load8<T>(reinterpret_cast<const T*>(p.src) + (((size_t)n * p.src_planes + p.src_plane0 + pl) * hw + px) * 8, v);
    load8<T>(reinterpret_cast<const T*>(p.res) + (((size_t)n * p.res_planes + p.res_plane0 + pl) * hw + px) * 8, r);
    const float* g = p.gate + (size_t)n * p.channels + pl * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], g[k], r[k]);
    store8<T>(reinterpret_cast<T*>(p.dst) + (((size_t)n * p.dst_planes + p.dst_plane0 + pl) * hw + px) * 8, v);
  }
}

// dst = src * w0[c] (+ src2): the per-channel gamma1 of MetaGated's `glob(x) * gamma1 + x` (arch.py:665)
template <typename T>
__global__ void __launch_bounds__(256) chan_affine_kernel(const __grid_constant__ TokenOpParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const int planes = p.channels >> 3;
  const size_t total = (size_t)p.n * planes * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t px = i % hw;
    const int pl = (int)((i / hw) % planes), n = (int)(i / (hw * planes));
    float v[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8<T>(reinterpret_cast<const T*>(p.src) + (((size_t)n * p.src_planes + p.src_plane0 + pl) * hw + px) * 8, v);
    if (p.src2 != nullptr) load8<T>(reinterpret_cast<const T*>(p.src2) + (((size_t)n * p.src2_planes + p.src2_plane0 + pl) * hw + px) * 8, r);
    const float4 ga = *reinterpret_cast<const float4*>(p.w0 + pl * 8), gb = *reinterpret_cast<const float4*>(p.w0 + pl * 8 + 4);
    v[0] = fmaf(v[0], ga.x, r[0]), v[1] = fmaf(v[1], ga.y, r[1]), v[2] = fmaf(v[2], ga.z, r[2]), v[3] = fmaf(v[3], ga.w, r[3]);
    v[4] = fmaf(v[4], gb.x, r[4]), v[5] = fmaf(v[5], gb.y, r[5]), v[6] = fmaf(v[6], gb.z, r[6]), v[7] = fmaf(v[7], gb.w, r[7]);
    store8<T>(reinterpret_cast<T*>(p.dst) + (((size_t)n * p.dst_planes + p.dst_plane0 + pl) * hw + px) * 8, v);
  }
}

// Raw LayerNorm sums {sum, sum of squares, 0, 0} per pixel of a bf16 map: what conv_tc's epilogue writes with Epi::ln_out, for the
// forwards that run a plan's convs on the CUDA-core kernel instead (the device-side cross-check mode)
__global__ void __launch_bounds__(256) ln_raw_sums_kernel(const __nv_bfloat16* src, int planes, int plane0, int nplanes, int n, size_t hw, float4* out,
                                                          int out_planes) {
  const size_t total = (size_t)n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / hw, px = i - b * hw;
    float s1 = 0.0f, s2 = 0.0f;
    for (int pl = 0; pl < nplanes; ++pl) {
      float v[8];
      load8<__nv_bfloat16>(src + ((b * planes + plane0 + pl) * hw + px) * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s1 += v[k], s2 = fmaf(v[k], v[k], s2);
    }
    out[b * out_planes * hw + px] = make_float4(s1, s2, 0.0f, 0.0f);
  }
}

}  // namespace

cudaError_t launch_ln_raw_sums(const void* src, int planes, int plane0, int nplanes, int n, int H, int W, void* out, int out_planes, int num_sms,
                               cudaStream_t s) {
  const size_t hw = (size_t)H * W;
  ln_raw_sums_kernel<<<grid_for((size_t)n * hw, 256, (num_sms > 0 ? num_sms : 148) * 16), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), planes, plane0, nplanes, n, hw, reinterpret_cast<float4*>(out), out_planes);
  return cudaGetLastError();
}

cudaError_t launch_chan_gate(const SeParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int sms = num_sms > 0 ? num_sms : 148;
  const dim3 g1(p.blocks, p.channels >> 3, p.n);
  if (bf16)
    se_pool_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>(p);
  else
    se_pool_kernel<float><<<g1, 256, 0, s>>>(p);
  sca_gate_kernel<<<p.n, 256, (size_t)p.channels * sizeof(float), s>>>(p);
  const int g = grid_for((size_t)p.n * (p.channels >> 3) * p.H * p.W, 256, sms * 32);
  if (bf16)
    sca_apply_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    sca_apply_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_chan_affine(const TokenOpParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int g = grid_for((size_t)p.n * (p.channels >> 3) * p.H * p.W, 256, (num_sms > 0 ? num_sms : 148) * 32);
  if (bf16)
    chan_affine_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    chan_affine_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_rmsnorm(const TokenOpParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int g = grid_for((size_t)p.n * p.H * p.W, 256, (num_sms > 0 ? num_sms : 148) * 32);
  if (bf16)
    rmsnorm_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    rmsnorm_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_unshuffle_pool(const TokenOpParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int g = grid_for((size_t)p.n * (p.channels >> 3) * p.H * p.W, 256, (num_sms > 0 ? num_sms : 148) * 32);
  if (bf16)
    unshuffle_pool_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    unshuffle_pool_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_dwconv_k(const TokenOpParams& p, int K, bool bf16, int num_sms, cudaStream_t s) {
  const int planes = (p.channels + 7) >> 3;
  const int gx = grid_for((size_t)p.H * p.W, 256, std::max(1, (num_sms > 0 ? num_sms : 148) * 16 / std::max(1, planes * p.n)));
  const dim3 g(gx, planes, p.n);
  if (K == 5 && p.dense5 && (long long)planes * p.n <= 65535) {
    const dim3 g5((p.W + 127) / 128, (p.H + kDw5Rows - 1) / kDw5Rows, planes * p.n);
    if (bf16)
      dwconv5_rows_kernel<__nv_bfloat16><<<g5, 128, 0, s>>>(p);
    else
      dwconv5_rows_kernel<float><<<g5, 128, 0, s>>>(p);
    return cudaGetLastError();
  }
  if (bf16)
    dwconv_k_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p, K);
  else
    dwconv_k_kernel<float><<<g, 256, 0, s>>>(p, K);
  return cudaGetLastError();
}

cudaError_t launch_se_shuffle(const SeParams& p, bool bf16, int num_sms, cudaStream_t s) {
  const int sms = num_sms > 0 ? num_sms : 148;
  if (p.gate != nullptr) {
    const dim3 g1(p.blocks, p.channels >> 3, p.n);
    if (bf16)
      se_pool_kernel<__nv_bfloat16><<<g1, 256, 0, s>>>(p);
    else
      se_pool_kernel<float><<<g1, 256, 0, s>>>(p);
    se_gate_kernel<<<p.n, 256, (size_t)(p.channels + p.hidden) * sizeof(float), s>>>(p);
  }
  const int g = grid_for((size_t)p.n * (p.channels >> 5) * p.H * p.W, 256, sms * 32);
  if (bf16)
    se_shuffle_kernel<__nv_bfloat16><<<g, 256, 0, s>>>(p);
  else
    se_shuffle_kernel<float><<<g, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace rsb
