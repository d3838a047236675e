// conv_direct — CUDA-core (FFMA, fp32 accumulate) convolution used for
//   * the whole fp32 plan (max-abs <= 1e-4 against the fp32 oracle is out of reach for tf32 tensor cores),
//   * layers the tensor-core kernel does not take (the first conv reading the caller's NCHW tensor, tiny Cin),
//   * a device-side cross-check of the tensor-core kernel (force_direct).
// Thread = one output pixel x 32 output channels; block = 32 x 8 pixel tile; weights streamed through shared
// memory one (input-plane, kernel-row) slab at a time so any Cin / kernel size fits.
// Also here: GroupNorm (+affine +skip) and the planar-8 -> NCHW debug read-back.
#include <algorithm>

#include "kernels.cuh"

namespace rsb {
namespace {

constexpr int kDW = 32, kDH = 8, kDThreads = kDW * kDH, kCoutGroup = 32;

template <typename T>
__device__ __forceinline__ void load_src8(const ConvDirectParams& p, int n, int plane, int sy, int sx, float (&o)[8]) {
  if (p.src_external) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = plane * 8 + i;
      o[i] = c < p.cin ? (ld_any(p.src, p.src_dtype, (((size_t)n * p.cin + c) * p.H + sy) * p.W + sx) - p.in_mean[c & 3]) * p.in_scale
                       : 0.0f;
    }
  } else if (p.src_upsample2) {
    load8<T>(reinterpret_cast<const T*>(p.src) + planar_index(n, p.src_planes, p.src_plane0 + plane, p.H >> 1, p.W >> 1, sy >> 1, sx >> 1), o);
  } else {
    load8<T>(reinterpret_cast<const T*>(p.src) + planar_index(n, p.src_planes, p.src_plane0 + plane, p.H, p.W, sy, sx), o);
  }
}

template <typename T, bool kFast>
__global__ void __launch_bounds__(kDThreads) conv_direct_kernel(const __grid_constant__ ConvDirectParams p) {
  extern __shared__ float wsm[];  // [kw][8][32]
  const int tiles_x = (p.W + kDW - 1) / kDW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int group = blockIdx.y, n = blockIdx.z;
  const int x = tx * kDW + (threadIdx.x & (kDW - 1));
  const int y = ty * kDH + (threadIdx.x / kDW);
  const bool valid = (x < p.W) && (y < p.H);

  float acc[kCoutGroup];
#pragma unroll
  for (int j = 0; j < kCoutGroup; ++j) acc[j] = 0.0f;

  const int slab = p.kw * 8 * kCoutGroup;
  for (int plane = 0; plane < p.cin_planes; ++plane) {
    for (int ky = 0; ky < p.kh; ++ky) {
      __syncthreads();
      const float* wsrc = p.wpack + (((size_t)plane * p.kh + ky) * p.kw) * 8 * p.cpad + group * kCoutGroup;
      for (int i = threadIdx.x; i < slab; i += kDThreads) {
        const int j = i & (kCoutGroup - 1), rc = i / kCoutGroup;  // rc = kx * 8 + c
        wsm[i] = wsrc[(size_t)rc * p.cpad + j];
      }
      __syncthreads();
      const int sy = y + ky - p.pad_t;
      if (!valid || sy < 0 || sy >= p.H) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int sx = x + kx - p.pad_l;
        if (sx < 0 || sx >= p.W) continue;
        float xin[8];
        load_src8<T>(p, n, plane, sy, sx, xin);
        const float4* w4 = reinterpret_cast<const float4*>(wsm + kx * 8 * kCoutGroup);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int j4 = 0; j4 < kCoutGroup / 4; ++j4) {
            const float4 w = w4[c * (kCoutGroup / 4) + j4];
            acc[4 * j4 + 0] = fmaf(xin[c], w.x, acc[4 * j4 + 0]);
            acc[4 * j4 + 1] = fmaf(xin[c], w.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(xin[c], w.z, acc[4 * j4 + 2]);
            acc[4 * j4 + 3] = fmaf(xin[c], w.w, acc[4 * j4 + 3]);
          }
        }
      }
    }
  }
  if (!valid) return;
  const int cstore = (p.epi.cout + 7) & ~7;
#pragma unroll
  for (int j8 = 0; j8 < kCoutGroup / 8; ++j8) {
    const int c0 = group * kCoutGroup + j8 * 8;
    if (c0 < cstore) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = acc[j8 * 8 + i];
      epilogue8<T, kFast>(p.epi, p.epi.bias, p.epi.slopes, v, c0, n, y, x);
    }
  }
}

// ---------------------------------------------------------------- im2col pack of the external input
// generic version: one thread per (pixel, 8-channel plane)
__global__ void __launch_bounds__(256) pack_input_kernel(const __grid_constant__ PackParams p) {
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = (size_t)p.n * p.kplanes * hw;
  const int kreal = p.cin * p.kh * p.kw;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.W);
    const int y = (int)((i / p.W) % p.H);
    const int plane = (int)((i / hw) % p.kplanes);
    const int n = (int)(i / (hw * p.kplanes));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = plane * 8 + j;
      float val = 0.0f;
      if (k < kreal) {
        const int ci = k / (p.kh * p.kw), t = k - ci * p.kh * p.kw;
        const int ky = t / p.kw, kx = t - ky * p.kw;
        const int sy = y + ky - p.pad_t, sx = x + kx - p.pad_l;
        if (sy >= 0 && sy < p.H && sx >= 0 && sx < p.W)
          val = (ld_any(p.src, p.src_dtype, (((size_t)n * p.cin + ci) * p.H + sy) * p.W + sx) - p.in_mean[ci & 3]) * p.in_scale;
      }
      v[j] = val;
    }
    store8<__nv_bfloat16>(dst + i * 8, v);
  }
}

// compile-time geometry (the RGB 3x3 stem every in-scope model starts with): one thread per pixel gathers its whole
// receptive field into registers (all indices static) and writes every plane
template <typename TIn, int CIN, int KH, int KW>
__global__ void __launch_bounds__(256) pack_input_fixed_kernel(const __grid_constant__ PackParams p) {
  constexpr int KREAL = CIN * KH * KW, KPLANES = (KREAL + 15) / 16 * 2;
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = (size_t)p.n * hw;
  const TIn* src = reinterpret_cast<const TIn*>(p.src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.W);
    const int y = (int)((i / p.W) % p.H);
    const int n = (int)(i / hw);
    float v[KPLANES * 8];
#pragma unroll
    for (int k = 0; k < KPLANES * 8; ++k) v[k] = 0.0f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const TIn* plane = src + ((size_t)n * CIN + ci) * hw;
      const float mean = p.in_mean[ci & 3];
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int sy = y + ky - KH / 2;
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const int sx = x + kx - KW / 2;
          if (sy >= 0 && sy < p.H && sx >= 0 && sx < p.W)
            v[(ci * KH + ky) * KW + kx] = ((float)plane[(size_t)sy * p.W + sx] - mean) * p.in_scale;
        }
      }
    }
#pragma unroll
    for (int pl = 0; pl < KPLANES; ++pl) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[pl * 8 + j];
      store8<__nv_bfloat16>(dst + (((size_t)n * KPLANES + pl) * hw + (size_t)y * p.W + x) * 8, o);
    }
  }
}

// ---------------------------------------------------------------- GroupNorm
// pass 1: per (n, group, block) partial sum / sum of squares in fp64 (deterministic two-stage reduce, no atomics)
template <typename T>
__global__ void __launch_bounds__(256) groupnorm_partial_kernel(const __grid_constant__ GroupNormParams p) {
  const int g = blockIdx.y, n = blockIdx.z;
  const int planes_per_group = p.channels / p.groups / 8;
  const size_t hw = (size_t)p.H * p.W;
  const size_t total = hw * planes_per_group;  // 8-channel chunks in this group
  double s = 0.0, ss = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pl = (int)(i / hw);
    const size_t pix = i - (size_t)pl * hw;
    float v[8];
    load8<T>(reinterpret_cast<const T*>(p.src) + (((size_t)n * p.src_planes + p.src_plane0 + g * planes_per_group + pl) * hw + pix) * 8, v);
    // fp32 inside one 8-channel chunk, fp64 across chunks: 2 double adds per 8 elements instead of 16 (fp64 issue was half
    // of this kernel's time); still a fixed summation order
    float s8 = 0.0f, ss8 = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s8 += v[k];
      ss8 = fmaf(v[k], v[k], ss8);
    }
    s += (double)s8;
    ss += (double)ss8;
  }
  __shared__ double sh[2][256];
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double* out = p.partial + (((size_t)n * p.groups + g) * p.blocks_per_group + blockIdx.x) * 2;
    out[0] = sh[0][0];
    out[1] = sh[1][0];
  }
}

// pass 2: every block re-reduces the (few) partials in a fixed order, then normalises + affine + skip
template <typename T>
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const __grid_constant__ GroupNormParams p) {
  const int g = blockIdx.y, n = blockIdx.z;
  const int cpg = p.channels / p.groups;
  const int planes_per_group = cpg / 8;
  const size_t hw = (size_t)p.H * p.W;
  __shared__ float s_mean, s_rstd;
  __shared__ double red[2][256];
  {
    // all 256 threads fetch partials at once and reduce them as a fixed tree (one thread walking them serially cost every
    // block ~25 us of dependent L2 round trips before its first load)
    const double* in = p.partial + ((size_t)n * p.groups + g) * p.blocks_per_group * 2;
    double s = 0.0, ss = 0.0;
    for (int b = threadIdx.x; b < p.blocks_per_group; b += blockDim.x) {
      s += in[2 * b];
      ss += in[2 * b + 1];
    }
    red[0][threadIdx.x] = s;
    red[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        red[0][threadIdx.x] += red[0][threadIdx.x + o];
        red[1][threadIdx.x] += red[1][threadIdx.x + o];
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const double s = red[0][0], ss = red[1][0];
    const double cnt = (double)hw * cpg;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean = (float)mean;
    s_rstd = (float)(1.0 / sqrt(var + (double)p.eps));
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  const size_t total = hw * planes_per_group;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pl = (int)(i / hw);
    const size_t pix = i - (size_t)pl * hw;
    const int plane = g * planes_per_group + pl;
    float v[8];
    load8<T>(reinterpret_cast<const T*>(p.src) + (((size_t)n * p.src_planes + p.src_plane0 + plane) * hw + pix) * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * p.gamma[plane * 8 + k] + p.beta[plane * 8 + k];
    if (p.skip != nullptr) {
      float r[8];
      load8<T>(reinterpret_cast<const T*>(p.skip) + (((size_t)n * p.skip_planes + p.skip_plane0 + plane) * hw + pix) * 8, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    }
    store8<T>(reinterpret_cast<T*>(p.dst) + (((size_t)n * p.dst_planes + p.dst_plane0 + plane) * hw + pix) * 8, v);
  }
}

template <typename T>
__global__ void planar_to_nchw_kernel(const T* src, int n, int planes, int plane0, int channels, int H, int W, float* dst) {
  const size_t total = (size_t)n * channels * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((size_t)W * H)) % channels);
    const int b = (int)(i / ((size_t)W * H * channels));
    dst[i] = (float)src[planar_index(b, planes, plane0 + (c >> 3), H, W, y, x) + (c & 7)];
  }
}

}  // namespace

cudaError_t launch_conv_direct(const ConvDirectParams& p, bool bf16_storage, cudaStream_t stream) {
  const dim3 grid(((p.W + kDW - 1) / kDW) * ((p.H + kDH - 1) / kDH), p.cpad / kCoutGroup, p.n);
  const size_t smem = (size_t)p.kw * 8 * kCoutGroup * sizeof(float);
  if (bf16_storage)
    conv_direct_kernel<__nv_bfloat16, true><<<grid, kDThreads, smem, stream>>>(p);
  else
    conv_direct_kernel<float, false><<<grid, kDThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_pack_input(const PackParams& p, cudaStream_t stream) {
  const size_t total = (size_t)p.n * p.kplanes * p.H * p.W;
  if (p.cin == 3 && p.kh == 3 && p.kw == 3) {
    const size_t pixels = (size_t)p.n * p.H * p.W;
    const int g = (int)std::min<size_t>((pixels + 255) / 256, 148 * 32);
    if (p.src_dtype == RSB_BF16)
      pack_input_fixed_kernel<__nv_bfloat16, 3, 3, 3><<<g, 256, 0, stream>>>(p);
    else if (p.src_dtype == RSB_F16)
      pack_input_fixed_kernel<__half, 3, 3, 3><<<g, 256, 0, stream>>>(p);
    else
      pack_input_fixed_kernel<float, 3, 3, 3><<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
  }
  if (p.cin == 3 && p.kh == 1 && p.kw == 1 && p.kplanes == 2) {  // NCHW RGB -> planar-8 with 16 channels (13 of them zero)
    const size_t pixels = (size_t)p.n * p.H * p.W;
    const int g = (int)std::min<size_t>((pixels + 255) / 256, 148 * 32);
    if (p.src_dtype == RSB_BF16)
      pack_input_fixed_kernel<__nv_bfloat16, 3, 1, 1><<<g, 256, 0, stream>>>(p);
    else if (p.src_dtype == RSB_F16)
      pack_input_fixed_kernel<__half, 3, 1, 1><<<g, 256, 0, stream>>>(p);
    else
      pack_input_fixed_kernel<float, 3, 1, 1><<<g, 256, 0, stream>>>(p);
    return cudaGetLastError();
  }
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  pack_input_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_groupnorm(const GroupNormParams& p, bool bf16_storage, cudaStream_t stream) {
  const dim3 grid(p.blocks_per_group, p.groups, p.n);
  if (bf16_storage) {
    groupnorm_partial_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
    groupnorm_apply_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  } else {
    groupnorm_partial_kernel<float><<<grid, 256, 0, stream>>>(p);
    groupnorm_apply_kernel<float><<<grid, 256, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_planar_to_nchw(const void* src, bool bf16_storage, int n, int planes, int plane0, int channels, int H,
                                  int W, float* dst, cudaStream_t stream) {
  const size_t total = (size_t)n * channels * H * W;
  const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  if (bf16_storage)
    planar_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), n, planes, plane0, channels, H, W, dst);
  else
    planar_to_nchw_kernel<float><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(src), n, planes, plane0, channels, H, W, dst);
  return cudaGetLastError();
}

}  // namespace rsb
