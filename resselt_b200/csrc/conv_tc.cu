// conv_tc — stride-1 'same' convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels x Cout] = sum over taps (dy,dx) and 16-channel K steps of
//                          A_tap[128 pixels x 16 ch] * W_tap[16 ch x Cout]        (tcgen05.mma, kind::f16, bf16 in, fp32 acc)
//
// One CTA per SM, persistent over 16x8-pixel output tiles, warp-specialised:
//   warp 0   : TMA producer. One 4-D box per tile brings the (16+kh-1) x (8+kw-1) halo tile of every input plane
//              into shared memory; out-of-image coordinates are zero-filled by TMA == the conv's zero padding.
//              Because activations are stored planar-8 ([C/8][H][W][8]), the box lands as [plane][row][pixel][8ch]:
//              8 consecutive pixels x 16 B are exactly one canonical no-swizzle K-major core matrix. A tap is then
//              just a different start address inside the same halo tile (+ (dy*WT + dx) * 16 B): the 9 taps of a 3x3
//              re-use one shared-memory copy, no im2col, no per-tap reload.
//   warp 1   : single-thread tcgen05.mma issuer; accumulators live in TMEM, num_acc (<= 4) buffers deep.
//   warp 2   : TMEM allocation / release.
//   warps 4+ : num_acc epilogue warpgroups; warpgroup g drains accumulator g (tiles g, g+A, g+2A, ...), so up to four
//              tile epilogues are in flight per SM and hide each other's TMEM / global-load / MUFU latencies while
//              the MMA warp runs ahead.  tcgen05.ld (32 lanes x 16 columns) -> bias / activation / gate / residual /
//              PixelShuffle (kernels.cuh::epilogue8) -> 16-byte stores, 8 neighbouring pixels filling a 128-byte line.
// The packed weights of the layer ([tap][cin/8][npad][8] bf16, also canonical K-major), its bias and PReLU slopes
// stay resident in shared memory for the CTA's lifetime.
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"

namespace rsb {

namespace {

constexpr int kMaxAcc = 4;
constexpr int kMaxThreads = 128 + 128 * kMaxAcc;
constexpr uint32_t kAlign = 1024;

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// KH/KW/KSTEPS > 0: geometry known at compile time -> the MMA issue loop is straight-line code (one UIADD3.64 + one
// UTCHMMA per MMA); KH == 0: geometry read from the parameter block (nested runtime loops).
// NCH > 0: the accumulator is NCH 16-column chunks wide (npad == 16 * NCH) -> unrolled epilogue with the residual
// operand prefetched before the accumulator-ready wait.
template <int KH, int KW, int KSTEPS, int NCH, int ACT, int COMB, int EXT>
__global__ void __launch_bounds__(kMaxThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using namespace ptx;
  using T = __nv_bfloat16;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const int A = p.num_acc;
  pdl_launch_dependents();
  // N-split group (ConvTcParams::nsplit): this CTA's slice, and its position among the CTAs of that slice
  const bool split = p.nsplit > 1;
  const int sl = split ? (int)blockIdx.x % p.nsplit : 0;
  const int bid = split ? (int)blockIdx.x / p.nsplit : (int)blockIdx.x;
  const int nblk = split ? (int)gridDim.x / p.nsplit : (int)gridDim.x;
  const void* const wpack_g = split ? p.slice[sl].wpack : p.wpack;
  const float* const bias_g = split ? p.slice[sl].bias : p.epi.bias;
  const float* const aux_g = split ? p.slice[sl].aux : (p.epi.ln_stats != nullptr ? p.epi.ln_rowsum : p.epi.slopes);
  const int coff = split ? p.slice[sl].dst_plane_off * 8 : 0;  // channel offset of the slice inside the destination range
  const int cout_s = split ? p.slice[sl].cout : p.epi.cout;

  const uint32_t w_al = align_up(p.wbytes, kAlign);
  const uint32_t st_al = align_up(p.stage_bytes, kAlign);
  uint8_t* const wsm = smem;
  uint8_t* const stage0 = smem + w_al;
  float* const bias_sm = reinterpret_cast<float*>(stage0 + (size_t)S * st_al);
  float* const slope_sm = bias_sm + p.npad;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(slope_sm + p.npad);
  uint64_t* const full = bars;
  uint64_t* const empty = bars + S;
  uint64_t* const tfull = bars + 2 * S;
  uint64_t* const tempty = tfull + kMaxAcc;
  uint64_t* const wbar = tempty + kMaxAcc;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  // seen[s] = number of fills of stage s whose completion an MMA warp has observed.  mbarrier waits only know the phase
  // PARITY: a warp that asks for fill k of a stage while fill k-1 has not completed yet is told "done" (it sees fill k-2).
  // With two issuing warps the previous fill of a stage may be the OTHER warp's tile, and TMA loads complete out of order
  // (L2 hit vs miss), so before waiting for fill k a warp first waits until whoever owned fill k-1 has seen it land.
  volatile uint32_t* const seen = tmem_slot + 4;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      seen[s] = 0;
    }
    for (int a = 0; a < kMaxAcc; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4 * (kMaxAcc / A));  // every epilogue warp that reads accumulator a arrives once
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.npad; i += blockDim.x) {
    bias_sm[i] = bias_g[i];
    // (a conv with a folded LayerNorm has no PReLU: the slot holds the weights' row sums instead)
    slope_sm[i] = aux_g != nullptr ? aux_g[i] : 0.0f;
  }
  const int HT = kTileH + p.kh - 1;
  const int WT = kTileW + p.kw - 1;
  const uint32_t a_lbo = (uint32_t)(HT * WT) * 16u;  // next 8-channel plane
  const uint32_t a_sbo = (uint32_t)WT * 16u;         // next tile row (8 pixels = one core matrix)
  const uint32_t b_lbo = (uint32_t)p.npad * 16u;     // next 8-input-channel slab
  const uint32_t b_sbo = 128u;                       // next 8 output channels
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything below reads or writes activation buffers

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&src_map);
      mbar_expect_tx(wbar, p.wbytes);
      for (uint32_t off = 0; off < p.wbytes; off += 32768u) {
        const uint32_t len = min(32768u, p.wbytes - off);
        bulk_load_1d(wsm + off, reinterpret_cast<const uint8_t*>(wpack_g) + off, len, wbar);
      }
      int s = 0;
      uint32_t ph = 0;
      const int cp8 = p.kchunk >> 3;  // planes per K chunk (one shared-memory stage holds one chunk of one tile)
      for (int tile = bid; tile < p.num_tiles; tile += nblk) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait_parked(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], p.stage_bytes);
          tma_load_4d(stage0 + (size_t)s * st_al, &src_map, &full[s], 8 * (tx * kTileW - p.pad_l),
                      ty * kTileH - p.pad_t, p.src_plane0 + c * cp8, n);
          if (++s == S) s = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // Two issuing warps (warp 1: even tiles, warp 3: odd tiles): a single thread cannot generate descriptors fast
    // enough to keep the tensor pipe full with N = 48 MMAs (~24 math cycles each).  Each tile's MMAs stay in order
    // inside one warp, so the accumulation order per output pixel is fixed.  The whole warp walks the loop (so every
    // operand stays in uniform registers); one elected lane issues.
    // K-chunked tiles share one ring of stages in chunk order; the seen[] protocol below keeps the parity waits exact for
    // both warps.  solo (bring-up switch): warp 1 alone issues every tile.
    const bool solo = p.solo_issue != 0;
    const int first = (warp == 1 || solo) ? 0 : 1;
    const int tstep = solo ? 1 : 2;
    const bool leader = elect_one();
    mbar_wait(wbar, 0);
    const uint32_t idesc = make_idesc_bf16(128, p.npad);
    const uint32_t a_hi = (uint32_t)(make_smem_desc(0, a_lbo, a_sbo) >> 32);
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, b_lbo, b_sbo) >> 32);
    const uint32_t a_lo_const = (uint32_t)make_smem_desc(0, a_lbo, a_sbo);  // LBO field of the low word
    const uint32_t b_lo0 = (uint32_t)make_smem_desc(smem_u32(wsm), b_lbo, b_sbo);
    const uint32_t a_kstep = 2u * (a_lbo >> 4);  // descriptor address units (16 B) per 16-channel K step
    const uint32_t b_kstep = 2u * (b_lbo >> 4);
    int i = first;
    for (int tile = bid + first * nblk; tile < p.num_tiles && !(solo && warp == 3); tile += tstep * nblk, i += tstep) {
      const int acc = i % A;
      const uint32_t aph = (uint32_t)(i / A) & 1u;
      mbar_wait(&tempty[acc], aph ^ 1);
      const uint32_t d = tmem_base + (uint32_t)acc * p.acc_stride;
      if constexpr (KH > 0) {
        const int s = i % S;  // static geometry: one chunk per tile
        const uint32_t fill = (uint32_t)(i / S);
        while (seen[s] < fill) {
        }
        mbar_wait(&full[s], fill & 1u);
        if (leader) seen[s] = fill + 1;
        tc_fence_after();
        constexpr uint32_t kPlane = (uint32_t)((kTileH + KH - 1) * (kTileW + KW - 1));  // 16-byte units per 8-ch plane
        const uint64_t da = make_smem_desc(smem_u32(stage0 + (size_t)s * st_al), kPlane * 16u, (uint32_t)(kTileW + KW - 1) * 16u);
        const uint64_t db = make_smem_desc(smem_u32(wsm), b_lbo, b_sbo);
        if (leader) {
#pragma unroll
          for (int dy = 0; dy < KH; ++dy)
#pragma unroll
            for (int dx = 0; dx < KW; ++dx)
#pragma unroll
              for (int kk = 0; kk < KSTEPS; ++kk) {
                const uint64_t a = da + (uint64_t)((uint32_t)(dy * (kTileW + KW - 1) + dx) + 2u * kk * kPlane);
                const uint64_t b = db + (uint64_t)((uint32_t)((dy * KW + dx) * KSTEPS + kk) * b_kstep);
                umma_bf16(d, a, b, idesc, (dy | dx | kk) != 0 ? 1u : 0u);
              }
        }
        if (leader) umma_commit(&empty[s]);  // shared-memory stage may be refilled once these MMAs have read it
      } else {
        // runtime geometry, K-chunked: chunk c of this tile sits in stage (i * nchunks + c) % S
        const int kc_steps = p.kchunk >> 4;
        const uint32_t cin8_units = (uint32_t)(p.cin >> 3) * (b_lbo >> 4);  // B descriptor units per tap
        uint32_t accum = 0;
        for (int c = 0; c < p.nchunks; ++c) {
          const int j = i * p.nchunks + c;
          const int s = j % S;
          const uint32_t fill = (uint32_t)(j / S);
          while (seen[s] < fill) {
          }
          mbar_wait(&full[s], fill & 1u);
          if (leader) seen[s] = fill + 1;
          tc_fence_after();
          const uint32_t a_lo0 = a_lo_const + (smem_u32(stage0 + (size_t)s * st_al) >> 4);
          uint32_t b_tap = b_lo0 + (uint32_t)(c * (p.kchunk >> 3)) * (b_lbo >> 4);
          for (int dy = 0; dy < p.kh; ++dy) {
            for (int dx = 0; dx < p.kw; ++dx) {
              uint32_t a_lo = a_lo0 + (uint32_t)(dy * WT + dx);
              uint32_t b_lo = b_tap;
              for (int kk = 0; kk < kc_steps; ++kk) {
                if (leader) umma_bf16_lohi_rt(d, a_lo, a_hi, b_lo, b_hi, idesc, accum);
                accum = 1;
                a_lo += a_kstep;
                b_lo += b_kstep;
              }
              b_tap += cin8_units;
            }
          }
          if (leader) umma_commit(&empty[s]);
        }
      }
      if (leader) umma_commit(&tfull[acc]);  // accumulator complete
    }
    __syncwarp();
  } else if (warp >= 4) {
    // four epilogue warpgroups share the A accumulator buffers: warpgroup wg drains accumulator wg % A and, when
    // A < 4 (wide N), only every (4/A)-th 16-column chunk of it, so wide layers still get four warpgroups of epilogue
    const int wg = (warp - 4) >> 2;
    const int g = wg % A, part = wg / A, parts = kMaxAcc / A;
    {
      const int q = warp & 3;  // TMEM lane quarter this warp may read
      const int row = q * 32 + lane;
      const int ry = row >> 3, rx = row & 7;
      const int cstore = (cout_s + 7) & ~7;
      // a slice of an N-split group addresses bias / destination by its channel offset inside the group's destination range
      const float* const bias_e = bias_sm - coff;
      const float* const slope_e = slope_sm - coff;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * p.acc_stride;
      uint32_t aph = 0;
      for (int tile = bid + g * nblk; tile < p.num_tiles; tile += A * nblk) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y = ty * kTileH + ry, x = tx * kTileW + rx;
        const bool valid = (y < p.H) && (x < p.W);
        if constexpr (NCH > 0) {
          constexpr bool kUsesRes = COMB == RSB_COMB_SPAB_GATE || COMB == RSB_COMB_MUL || COMB == RSB_COMB_AXPY;
          uint4 pre[kUsesRes ? 2 * NCH : 1];
          if constexpr (kUsesRes) {
            // the residual does not depend on this tile's MMAs: fetch it while they are still running
            if (valid) {
              const T* rp = reinterpret_cast<const T*>(p.epi.res1) + planar_index(n, p.epi.res1_planes, p.epi.res1_plane0, p.H, p.W, y, x);
              const size_t plane_stride = (size_t)p.H * p.W * 8;
#pragma unroll
              for (int j = 0; j < 2 * NCH; ++j)
                if (j * 8 < cstore) pre[j] = *reinterpret_cast<const uint4*>(rp + j * plane_stride);
            }
          }
          const bool ln = p.epi.ln_stats != nullptr;
          float2 lnst = make_float2(1.0f, 0.0f);
          if (ln && valid) lnst = ln_stats_of(p.epi, n, y, x);  // folded LayerNorm: {rstd, -mean * rstd} of this pixel
          mbar_wait_parked(&tfull[g], aph);
          tc_fence_after();
          // software-pipelined accumulator reads: chunk ci+1 is in flight while chunk ci goes through the epilogue
          uint32_t r[2][16];
          tmem_ld16(taddr, r[0]);
          const bool lean = EXT == 0 && p.epi.simple;
          T* const drow = reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0 + (coff >> 3), p.H, p.W, y, x);
          const size_t dstride = (size_t)p.H * p.W * 8;
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            const int c = ci * 16;
            tmem_ld_wait();
            if (ci + 1 < NCH) tmem_ld16(taddr + (uint32_t)(c + 16), r[(ci + 1) & 1]);
            if (lean) {
              if (valid) epilogue16_planar<ACT, COMB>(p.epi, bias_sm, slope_sm, r[ci & 1], c, cstore, drow, dstride, kUsesRes ? &pre[2 * ci] : nullptr);
            } else if (valid) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[ci & 1][j]);
              if (ln) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], lnst.x, lnst.y * slope_sm[c + j]);
              }
              if (c < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_e, slope_e, v, coff + c, n, y, x, kUsesRes ? &pre[2 * ci] : nullptr, true);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[ci & 1][8 + j]);
              if (ln) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], lnst.x, lnst.y * slope_sm[c + 8 + j]);
              }
              if (c + 8 < cstore)
                epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_e, slope_e, v, coff + c + 8, n, y, x, kUsesRes ? &pre[2 * ci + 1] : nullptr, true);
            }
          }
        } else {
          // wide / runtime N: the residual chunks of a gate / AXPY / MUL tail are loaded inside epilogue8, right where they are
          // needed (no registers to hold 12 of them): ncu showed 44 % of this kernel's stall samples on the first use of that load
          // (profiles/r2_conv_tc_axpy.md) — 12 serial HBM round trips per tile, 85 us for a 180 -> 180 linear with a residual
          // against 39 us without.  So the lines of the NEXT tile this warpgroup will drain are pulled into L2 now, one tile
          // (~5 us) ahead: the later loads then pay an L2 hit.
          constexpr bool kMayRes = COMB == kRuntime || COMB == RSB_COMB_SPAB_GATE || COMB == RSB_COMB_MUL || COMB == RSB_COMB_AXPY;
          if constexpr (kMayRes) {
            const int comb_rt = COMB == kRuntime ? p.epi.combine : COMB;
            if (comb_rt != RSB_COMB_NONE && !p.epi.dst_external && p.epi.dst_ps <= 1) {
              const size_t plane_stride = (size_t)p.H * p.W * 8;
              auto prefetch_tile = [&](int t) {
                if (t >= p.num_tiles) return;
                const int n2 = t / tiles_per_img, rem2 = t - n2 * tiles_per_img;
                const int ty2 = rem2 / p.tiles_x, tx2 = rem2 - ty2 * p.tiles_x;
                const int y2 = ty2 * kTileH + ry, x2 = tx2 * kTileW + rx;
                if (y2 >= p.H || x2 >= p.W) return;
                const T* rp = reinterpret_cast<const T*>(p.epi.res1) + planar_index(n2, p.epi.res1_planes, p.epi.res1_plane0, p.H, p.W, y2, x2);
                for (int c = part * 16; c < cstore; c += 16 * parts) {
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)(c >> 3) * plane_stride));
                  if (c + 8 < cstore) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)((c >> 3) + 1) * plane_stride));
                }
              };
              if (tile == bid + g * nblk) prefetch_tile(tile);  // first tile of this warpgroup
              prefetch_tile(tile + A * nblk);
            }
          }
          const bool ln = p.epi.ln_stats != nullptr;
          float2 lnst = make_float2(1.0f, 0.0f);
          if (ln && valid) lnst = ln_stats_of(p.epi, n, y, x);  // folded LayerNorm: {rstd, -mean * rstd} of this pixel
          // Lean form for planar destinations (no sub-pixel scatter / split / second residual / border bias): row pointers once per
          // tile, per 8 channels one fused multiply-add per element (bias and the LayerNorm shift merged), the residual chunk, a
          // 16-byte store.  The token linears (K = 192: twelve MMAs per tile) are bound by the ISSUE of this epilogue, not by HBM:
          // the generic epilogue8 re-derives the planar address and re-tests the destination kind for every 8 channels.
          const int comb_rt = COMB == kRuntime ? p.epi.combine : COMB;
          const bool lean = (EXT == 0 || !p.epi.dst_external) && p.epi.dst_ps <= 1 && p.epi.dst2 == nullptr && p.epi.res2 == nullptr &&
                            p.epi.border_bias == nullptr && comb_rt != RSB_COMB_SPAB_GATE;
          const size_t pstride = (size_t)p.H * p.W * 8;
          T* const drow = lean && valid ? reinterpret_cast<T*>(p.epi.dst) + planar_index(n, p.epi.dst_planes, p.epi.dst_plane0 + (coff >> 3), p.H, p.W, y, x) : nullptr;
          const T* const rrow = lean && valid && comb_rt != RSB_COMB_NONE
                                    ? reinterpret_cast<const T*>(p.epi.res1) + planar_index(n, p.epi.res1_planes, p.epi.res1_plane0 + (coff >> 3), p.H, p.W, y, x)
                                    : nullptr;
          mbar_wait_parked(&tfull[g], aph);
          tc_fence_after();
          const bool lnout = p.epi.ln_out != nullptr;  // partial LayerNorm sums of the stored values (Epi::ln_out)
          float ls1 = 0.0f, ls2 = 0.0f;
          if (lean) {
            auto emit = [&](const uint32_t (&r)[16], int c) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int c0 = c + 8 * half;
                if (c0 >= cstore) break;
                const uint32_t* const r8 = &r[8 * half];
                const float4 b0 = *reinterpret_cast<const float4*>(bias_sm + c0), b1 = *reinterpret_cast<const float4*>(bias_sm + c0 + 4);
                float t[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float v[8];
                if (ln) {
                  const float4 s0 = *reinterpret_cast<const float4*>(slope_sm + c0), s1 = *reinterpret_cast<const float4*>(slope_sm + c0 + 4);
                  const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r8[j]), lnst.x, fmaf(lnst.y, sv[j], t[j]));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r8[j]) + t[j];
                }
                const int act_rt = ACT == kRuntime ? p.epi.act : ACT;
                if (act_rt != RSB_ACT_NONE) {
                  float sl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                  if (act_rt == RSB_ACT_PRELU) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) sl[j] = slope_sm[c0 + j];
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = activate<true, ACT>(p.epi.act, v[j], p.epi.act_param, sl[j]);
                }
                if (comb_rt != RSB_COMB_NONE) {
                  float rr[8];
                  unpack8<__nv_bfloat16>(*reinterpret_cast<const uint4*>(rrow + (size_t)(c0 >> 3) * pstride), rr);
                  if (comb_rt == RSB_COMB_MUL) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] *= rr[j];
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(p.epi.alpha, v[j], p.epi.beta1 * rr[j]);
                  }
                }
                if (lnout) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) ls1 += v[j], ls2 = fmaf(v[j], v[j], ls2);
                }
                store8<T>(drow + (size_t)(c0 >> 3) * pstride, v);
              }
            };
            // software-pipelined accumulator reads, two chunks per round: the next chunk's tcgen05.ld is in flight while this one
            // goes through the epilogue
            const int cstep = 16 * parts;
            uint32_t ra[16], rbb[16];
            int c = part * 16;
            if (c < p.npad) tmem_ld16(taddr + (uint32_t)c, ra);
            for (; c < p.npad; c += 2 * cstep) {
              tmem_ld_wait();
              if (c + cstep < p.npad) tmem_ld16(taddr + (uint32_t)(c + cstep), rbb);
              if (valid) emit(ra, c);
              if (c + cstep < p.npad) {
                tmem_ld_wait();
                if (c + 2 * cstep < p.npad) tmem_ld16(taddr + (uint32_t)(c + 2 * cstep), ra);
                if (valid) emit(rbb, c + cstep);
              }
            }
            if (lnout && valid) {  // this warpgroup's pair of the pixel's chunk (parts <= 2: conv_tc_num_acc)
              float2* o = reinterpret_cast<float2*>(reinterpret_cast<char*>(p.epi.ln_out) +
                                                    (((size_t)n * p.epi.ln_out_planes * p.H + y) * p.W + x) * (size_t)p.epi.ln_out_stride);
              o[part] = make_float2(ls1, ls2);
              if (parts == 1) o[1] = make_float2(0.0f, 0.0f);
            }
          } else
          for (int c = part * 16; c < p.npad; c += 16 * parts) {
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)c, r);
            tmem_ld_wait();
            if (!valid) continue;
            {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
              if (ln) {
                const float4 s0 = *reinterpret_cast<const float4*>(slope_sm + c), s1 = *reinterpret_cast<const float4*>(slope_sm + c + 4);
                v[0] = fmaf(v[0], lnst.x, lnst.y * s0.x), v[1] = fmaf(v[1], lnst.x, lnst.y * s0.y), v[2] = fmaf(v[2], lnst.x, lnst.y * s0.z), v[3] = fmaf(v[3], lnst.x, lnst.y * s0.w);
                v[4] = fmaf(v[4], lnst.x, lnst.y * s1.x), v[5] = fmaf(v[5], lnst.x, lnst.y * s1.y), v[6] = fmaf(v[6], lnst.x, lnst.y * s1.z), v[7] = fmaf(v[7], lnst.x, lnst.y * s1.w);
              }
              if (c < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_e, slope_e, v, coff + c, n, y, x, nullptr, true);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 + j]);
              if (ln) {
                const float4 s0 = *reinterpret_cast<const float4*>(slope_sm + c + 8), s1 = *reinterpret_cast<const float4*>(slope_sm + c + 12);
                v[0] = fmaf(v[0], lnst.x, lnst.y * s0.x), v[1] = fmaf(v[1], lnst.x, lnst.y * s0.y), v[2] = fmaf(v[2], lnst.x, lnst.y * s0.z), v[3] = fmaf(v[3], lnst.x, lnst.y * s0.w);
                v[4] = fmaf(v[4], lnst.x, lnst.y * s1.x), v[5] = fmaf(v[5], lnst.x, lnst.y * s1.y), v[6] = fmaf(v[6], lnst.x, lnst.y * s1.z), v[7] = fmaf(v[7], lnst.x, lnst.y * s1.w);
              }
              if (c + 8 < cstore) epilogue8<T, true, ACT, COMB, EXT>(p.epi, bias_e, slope_e, v, coff + c + 8, n, y, x, nullptr, true);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[g]);
        aph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

typedef void (*KernelFn)(const CUtensorMap, const ConvTcParams);

struct Variant {
  int kh, kw, ksteps, nch, act, comb, ext;
  KernelFn fn;
};

#define RSB_V(KH, KW, KS, NCH, ACT, COMB) {KH, KW, KS, NCH, ACT, COMB, 0, conv_tc_kernel<KH, KW, KS, NCH, ACT, COMB, 0>}
#define RSB_X(KH, KW, KS, NCH, ACT, COMB) {KH, KW, KS, NCH, ACT, COMB, kRuntime, conv_tc_kernel<KH, KW, KS, NCH, ACT, COMB, kRuntime>}
// Specialisations for the (geometry, epilogue) pairs the in-scope architectures emit.  Lookup order: exact match,
// then runtime geometry with the specialised epilogue, then the fully runtime kernel.
const Variant kVariants[] = {
    // SPAN / SPANPlus (48 channels): stem 1x1 over the im2col buffer, SPAB convs, conv_cat, upsampler
    RSB_V(1, 1, 2, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(3, 3, 3, 3, RSB_ACT_SILU, RSB_COMB_NONE),
    RSB_V(3, 3, 3, 3, RSB_ACT_MISH, RSB_COMB_NONE),
    RSB_V(3, 3, 3, 3, RSB_ACT_NONE, RSB_COMB_SPAB_GATE),
    RSB_V(3, 3, 3, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(1, 1, 12, 3, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_X(3, 3, 3, 0, RSB_ACT_NONE, RSB_COMB_NONE),  // upsampler conv -> PixelShuffle store into the caller's tensor
    // Compact (64 channels, PReLU)
    RSB_V(1, 1, 2, 0, RSB_ACT_PRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 4, 0, RSB_ACT_PRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 4, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_X(3, 3, 4, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    // ESRGAN dense blocks (64 + 32k input channels) and RealPLKSR
    RSB_V(3, 3, 4, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 6, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 8, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 10, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(3, 3, 6, 0, RSB_ACT_NONE, RSB_COMB_AXPY),  // the two halves of the K-split 192->64 conv
    RSB_V(3, 3, 4, 0, RSB_ACT_NONE, RSB_COMB_AXPY),
    RSB_V(2, 2, 4, 0, RSB_ACT_LRELU, RSB_COMB_NONE),  // 2x2 phase kernels of the nearest-x2 upconv
    RSB_V(3, 3, 4, 0, RSB_ACT_MISH, RSB_COMB_NONE),
    RSB_V(3, 3, 8, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(3, 3, 4, 0, RSB_ACT_SIGMOID, RSB_COMB_MUL),
    RSB_V(1, 1, 4, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    // runtime geometry, specialised epilogue
    RSB_V(0, 0, 0, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_V(0, 0, 0, 0, RSB_ACT_LRELU, RSB_COMB_NONE),
    RSB_V(0, 0, 0, 0, RSB_ACT_NONE, RSB_COMB_AXPY),
    RSB_V(0, 0, 0, 0, RSB_ACT_GELU, RSB_COMB_NONE),
    RSB_V(0, 0, 0, 0, RSB_ACT_MISH, RSB_COMB_NONE),  // gated-CNN blocks of RTMoSR / GateRV3: mish(fc2(.)), mish(g) * cat(i, c), mish(fc2(.)) + x
    RSB_V(0, 0, 0, 0, RSB_ACT_MISH, RSB_COMB_MUL),
    RSB_V(0, 0, 0, 0, RSB_ACT_MISH, RSB_COMB_AXPY),
    RSB_V(1, 1, 12, 0, RSB_ACT_NONE, RSB_COMB_NONE),  // DAT token linears: 192 -> 180
    RSB_V(1, 1, 12, 0, RSB_ACT_GELU, RSB_COMB_NONE),
    RSB_V(1, 1, 12, 0, RSB_ACT_NONE, RSB_COMB_AXPY),
    // fully runtime
    RSB_X(0, 0, 0, 0, RSB_ACT_NONE, RSB_COMB_NONE),
    RSB_X(0, 0, 0, 0, kRuntime, kRuntime),
};
#undef RSB_V
#undef RSB_X
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

KernelFn pick(const ConvTcParams& p) {
  int act = p.epi.act;
  const int comb = p.epi.combine;
  if (comb == RSB_COMB_SPAB_GATE) act = RSB_ACT_NONE;  // the gate ignores `act`
  const int ks = p.cin >> 4;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < kNumVariants; ++i) {
      const Variant& v = kVariants[i];
      const bool geo = pass == 0 ? (v.kh == p.kh && v.kw == p.kw && v.ksteps == ks && p.nchunks == 1) : v.kh == 0;
      const bool width = v.nch == 0 || v.nch * 16 == p.npad;
      const bool ext = p.epi.dst_external ? v.ext == kRuntime : v.ext == 0;
      if (p.epi.ln_out != nullptr && v.nch != 0) continue;  // the partial LayerNorm sums live in the runtime-N lean epilogue
      if (geo && width && ext && v.act == act && v.comb == comb) return v.fn;
    }
  return kVariants[kNumVariants - 1].fn;
}

}  // namespace

size_t conv_tc_smem_bytes(int cin, int kchunk, int npad, int kh, int kw, int stages) {
  const uint32_t wbytes = (uint32_t)kh * kw * cin * npad * 2u;
  const uint32_t stage = (uint32_t)(kTileH + kh - 1) * (kTileW + kw - 1) * kchunk * 2u;
  return (size_t)align_up(wbytes, kAlign) + (size_t)stages * align_up(stage, kAlign) + 2 * npad * sizeof(float) +
         (2 * stages + 2 * kMaxAcc + 1) * 8 + 16 + 8 * sizeof(uint32_t);  // barriers, TMEM slot, seen[] (<= 8 stages)
}

int conv_tc_num_acc(int npad) {
  const int a = 512 / npad;
  return a >= 4 ? 4 : (a >= 2 ? 2 : 1);
}

cudaError_t conv_tc_configure(size_t max_smem) {
  for (int i = 0; i < kNumVariants; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kVariants[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_tc(const CUtensorMap& src_map, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem = conv_tc_smem_bytes(p.cin, p.kchunk, p.npad, p.kh, p.kw, p.stages);
  const int ns = p.nsplit > 1 ? p.nsplit : 1;
  const int per = std::max(1, num_sms / ns);  // CTAs per slice
  const int grid = (p.num_tiles < per ? p.num_tiles : per) * ns;
  const int threads = kMaxThreads;
  KernelFn fn = pick(p);
  return launch_pdl(fn, dim3(grid), dim3(threads), smem, stream, src_map, p);
}

}  // namespace rsb
