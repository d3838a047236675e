/*
 * resselt_b200 — C ABI of the B200 (sm_100a) super-resolution forward engine.
 *
 * This is the drop-in boundary for the hot path of rewaifu/resselt: the `forward` of the
 * architectures it loads.  The reference has no native layer; each entry point below replaces a
 * group of ATen calls issued from the reference's Python `nn.Module.forward`:
 *
 *   rsb_plan_add_conv    <- nn.Conv2d / F.conv2d call sites
 *                           (resselt/archs/span/arch.py:110-117,152-154,222;
 *                            resselt/archs/spanplus/arch.py:49-56,100,141,180-184;
 *                            resselt/archs/compact/arch.py:39,46,52;
 *                            resselt/utilities/block.py:176-185;
 *                            resselt/archs/plksr/rplksr.py:15,17,27,45,79,123,126)
 *                           with the element-wise tails that follow them fused in
 *                           (SiLU/Mish/PReLU/LeakyReLU/sigmoid: span/arch.py:169-177,
 *                            spanplus/arch.py:119-127, compact/arch.py:41,48, utilities/block.py:25;
 *                            SPAB gate span/arch.py:176-177; residual adds utilities/block.py:344,465;
 *                            PixelShuffle span/arch.py:55, compact/arch.py:54-64, rplksr.py:143-147).
 *   rsb_plan_add_groupnorm <- nn.GroupNorm + skip (resselt/archs/plksr/rplksr.py:83,91-93)
 *   rsb_plan_add_op      <- LayerNorm / depthwise conv / window + channel attention / AIM call sites of DAT and SwinIR, and the
 *                           DySample head of SPANPlus / RealPLKSR (resselt/utilities/dysample.py:46-83) (listed at rsb_op_kind below)
 *   rsb_plan_forward     <- <Module>.forward (span/arch.py:231-250, spanplus/arch.py:199-201,
 *                           compact/arch.py:56-65, esrgan/arch.py:129-138, plksr/rplksr.py:145-147,
 *                           dat/arch.py:970-990, swinir/arch.py:962-1011)
 *
 * Conventions: plain C types only; every function returns 0 on success, a negative rsb_status for
 * argument/state errors, or a positive cudaError_t value; rsb_last_error() returns a thread-local
 * message for the last failure.  The library never allocates activation memory: the caller passes
 * a workspace of rsb_plan_workspace_bytes() bytes.  Host weight pointers are copied during
 * rsb_plan_add_* and may be freed afterwards.  All work is enqueued on the caller's CUDA stream.
 */
#ifndef RESSELT_B200_H
#define RESSELT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSB_VERSION 204

typedef struct rsb_plan rsb_plan;

enum rsb_status {
  RSB_OK = 0,
  RSB_ERR_INVALID = -1,     /* bad argument */
  RSB_ERR_STATE = -2,       /* call not valid in the plan's current state */
  RSB_ERR_UNSUPPORTED = -3, /* shape / feature not implemented */
  RSB_ERR_WORKSPACE = -4,   /* workspace too small or misaligned */
  RSB_ERR_NO_DEVICE = -5    /* no usable sm_100 device / driver entry point */
};

enum rsb_dtype { RSB_F32 = 0, RSB_BF16 = 1, RSB_F16 = 2 };

enum rsb_act {
  RSB_ACT_NONE = 0,
  RSB_ACT_SILU = 1,
  RSB_ACT_MISH = 2,
  RSB_ACT_LRELU = 3,  /* slope = act_param */
  RSB_ACT_PRELU = 4,  /* per-channel slopes */
  RSB_ACT_SIGMOID = 5,
  RSB_ACT_GELU = 6    /* exact erf form */
};

/* What happens to a = act(conv + bias) before it is stored. r1/r2 are other plan buffers. */
enum rsb_combine {
  RSB_COMB_NONE = 0,      /* out = a                                                  */
  RSB_COMB_SPAB_GATE = 1, /* out = (v + r1) * (sigmoid(v) - 0.5), v = conv + bias      */
  RSB_COMB_MUL = 2,       /* out = a * r1                                              */
  RSB_COMB_AXPY = 3       /* out = alpha * a + beta1 * r1 [+ beta2 * r2]               */
};

#define RSB_EXTERNAL_INPUT (-1)  /* src: the caller's NCHW input tensor   */
#define RSB_EXTERNAL_OUTPUT (-2) /* dst: the caller's NCHW output tensor  */
#define RSB_NO_BUFFER (-3)

/* One convolution ('same' zero padding, stride 1) with its fused tail. */
typedef struct rsb_conv_desc {
  int32_t src_buf;    /* buffer id or RSB_EXTERNAL_INPUT */
  int32_t src_ch_off; /* first input channel inside the buffer (multiple of 8) */
  int32_t cin;
  int32_t dst_buf;    /* buffer id or RSB_EXTERNAL_OUTPUT */
  int32_t dst_ch_off; /* multiple of 8 */
  int32_t cout;
  int32_t kh, kw;           /* kernel extents (odd unless pad_t/pad_l are given) */
  const float* weight;      /* host, [cout][cin][kh][kw] */
  const float* bias;        /* host, [cout] or NULL */
  int32_t act;              /* rsb_act */
  float act_param;          /* LeakyReLU slope */
  const float* act_slopes;  /* host, [cout] PReLU slopes or NULL */
  int32_t combine;          /* rsb_combine */
  int32_t res1_buf, res1_ch_off;
  int32_t res2_buf, res2_ch_off; /* RSB_NO_BUFFER when unused */
  float alpha, beta1, beta2;
  /* external input only: x' = (x - in_mean[c]) * in_scale, applied before zero padding */
  float in_mean[4];
  float in_scale;
  /* external output only: PixelShuffle(ps) into NCHW [n][cout/ps^2][h*ps][w*ps];
   * add_base: out[c][h*ps+i][w*ps+j] += x[c][h][w] (nearest-upsampled raw input);
   * finally out = out * out_scale + out_mean[c]. */
  int32_t ps;
  int32_t add_base;
  float out_scale;
  float out_mean[4];
  /* source sampled through a nearest-neighbour x2 upsample (src buffer lives on the half-size grid) */
  int32_t src_upsample2;
  /* buffer destinations only.
   * dst_ps > 1: sub-pixel convolution — cout = dst_ps^2 * C, channel blocks are phase-major
   *   (channel (a*dst_ps + b)*C + c is output channel c of pixel (y*dst_ps + a, x*dst_ps + b)); dst_buf lives on the
   *   dst_ps-times finer grid and receives C channels.
   * dst2_buf != RSB_NO_BUFFER: output channels >= split_ch are written to dst2_buf starting at dst2_ch_off
   *   (split_ch multiple of 8); channels < split_ch go to dst_buf as usual. */
  int32_t dst_ps;
  int32_t dst2_buf, dst2_ch_off, split_ch;
  /* dst_phase >= 0 (with dst_ps > 1): this conv produces only phase dst_phase = a*dst_ps + b (cout = C). */
  int32_t dst_phase;
  /* explicit top/left zero padding; -1 selects the 'same' default kh/2, kw/2.  With explicit pads the kernel
   * extents may be even (used for the 2x2 phase kernels of a nearest-upsample + 3x3 conv). Output size == input size. */
  int32_t pad_t, pad_l;
  /* host, [16][cout] or NULL: extra bias for pixels on the border of the conv grid, indexed by the mask
   * (y == 0) | (y == H-1) << 1 | (x == 0) << 2 | (x == W-1) << 3 (row 0 unused), added with the bias before the activation.
   * This is what makes a 1x1 conv WITH BIAS followed by a zero-padded k x k conv collapse exactly into one k x k conv
   * (the taps that fall outside the image must not see the 1x1 conv's bias): SPAN's conv_cat + upsampler, span/arch.py:247-248. */
  const float* border_bias;
  /* LayerNorm folded into this conv (1 x 1, buffer source).  ln_fold != 0: ln_stats_buf names an 8-channel buffer written by
   * RSB_OP_LAYERNORM in statistics mode (i[0] = 1) over this conv's source; the caller passes weight' = weight * gamma and
   * bias' = bias + weight . beta, and the conv computes  rstd * (weight' . x) - mean * rstd * rowsum(weight') + bias'
   * == weight . LN(x) + bias, with rowsum taken over the weights as rounded to the plan's dtype, so the mean term cancels
   * exactly.  One read of x instead of LayerNorm's read + write + the conv's read (swinir/arch.py:268-332 norm1 -> qkv,
   * norm2 -> fc1; dat/arch.py:565-612).  A zero-initialised descriptor has no fold. */
  int32_t ln_fold;
  int32_t ln_stats_buf;
  /* The statistics without their own pass.  ln_out != 0 (bf16 plans; 1 x 1 conv into a planar buffer, no sub-pixel / split / second
   * residual / border bias / SPAB gate): this conv's epilogue also writes, per pixel, the partial sums {sum v, sum v^2} of the values
   * it stores (all cout channels) into the pixel chunks of the 8-channel buffer ln_out_buf — two float pairs per chunk, one per
   * epilogue warpgroup.  A consumer reads them with ln_fold == 2 (raw sums: mean and rstd are derived in its epilogue from the sums,
   * its own cin and ln_eps) instead of ln_fold == 1 (the {rstd, -mean * rstd} the statistics-mode LayerNorm op writes).  Used for
   * `x += proj(..)` -> norm2 -> fc1 and `x += fc2(..)` -> norm1 -> qkv (swinir/arch.py:296-335, dat/arch.py:636-683): one pass over
   * x less per LayerNorm. */
  int32_t ln_out;
  int32_t ln_out_buf;
  float ln_eps; /* epsilon of the LayerNorm whose raw sums are consumed (ln_fold == 2) */
} rsb_conv_desc;

/* GroupNorm over (channels/groups, H, W) per sample, affine, followed by "+ skip". */
typedef struct rsb_groupnorm_desc {
  int32_t src_buf, src_ch_off;
  int32_t dst_buf, dst_ch_off;
  int32_t channels, groups;
  float eps;
  const float* gamma; /* host [channels] */
  const float* beta;  /* host [channels] */
  int32_t skip_buf, skip_ch_off; /* RSB_NO_BUFFER: no skip */
} rsb_groupnorm_desc;

/* Token-wise / attention ops of the transformer architectures (DAT, SwinIR) and the DySample head.  A token is a pixel of a
 * planar buffer.
 * Call sites replaced: /root/reference/resselt/archs/dat/arch.py:48,636,672,897,924 (LayerNorm), :49,345,547
 * (depthwise conv), :224-267 + :456-482 (shifted-window attention), :565-589 (channel attention), :492-508 and
 * :594-607 (adaptive interaction module); /root/reference/resselt/archs/swinir/arch.py:253,262,299,334,878,956
 * (LayerNorm), :133-170 + :268-332 (W-MSA / SW-MSA: window partition, relative-position bias gathered through
 * relative_position_index, cyclic shift and mask, window reverse). */
enum rsb_op_kind {
  RSB_OP_LAYERNORM = 1, /* dst = LN_channels(src) * w[0] + w[1];  f[0] = eps.  i[0] = 1: statistics only — dst is an 8-channel
                           buffer whose pixel chunks receive {rstd, -mean * rstd} as two floats (rsb_conv_desc.ln_stats_buf);
                           i[0] = 2: the same for a channels-first RMSNorm x / (|x|_2 / sqrt(C) + eps): {1 / (rms + eps), 0}, so
                           that a 1x1 conv with ln_fold = 1 computes conv(RMSNorm(x) * scale + offset) (gaterv3/arch.py:511-524)   */
  RSB_OP_DWCONV3 = 2,   /* dst = act(dwconv3x3(src; w[0] = [C][9], w[1] = bias[C])) [* src2];  i[0] = rsb_act;
                           i[1] = K in {0, 3, 5, 7}: depthwise K x K instead (w[0] = [C][K*K]; K > 3: no act / gate)          */
  RSB_OP_WINATTN = 3,   /* src = [q | k | v]; i[0] heads, i[1] split_h, i[2] split_w, i[3] shifted, i[4] channel stride
                           between q, k and v (0: channels);
                           f[0] = qk scale; w[0] / w[1] = position-bias tables of the two branches
                           ([(2 sh - 1)(2 sw - 1)][heads / 2], offset index (dy + sh - 1)(2 sw - 1) + dx + sw - 1).
                           Branch 0 (first half of the channels / heads) uses split_h x split_w windows, branch 1 the
                           transposed shape; square windows (split_h == split_w) give plain Swin attention, the two
                           tables then being the two head halves of the learned relative_position_bias_table.
                           i[5] = 32 (bf16 plans, head_dim < 32, window sides multiples of 8 with 64 / 128 / 256 tokens): head g of
                           q, k, v and of dst starts at channel 32 g (the caller pads its qkv weights with zero rows and its
                           proj weights with zero columns); channels 32 g + head_dim .. 32 g + 31 of dst are written as zeros.
                           This layout runs on the tcgen05 / TMEM kernel (winattn_tc.cu); i[5] = 0: heads packed             */
  RSB_OP_CHANATTN = 4,  /* src = [q | k | v]; i[0] heads, i[1] q/k/v channel stride (0: channels); w[0] = temperature[heads] */
  RSB_OP_AIM = 5        /* src = attention output, src2 = conv branch, dst = gated sum; i[0] mode (0 window block,
                           1 channel block), i[1] / i[2] hidden widths of the channel / spatial MLPs;
                           w[0..3] = channel MLP (W1 [h1][C], b1, W2 [C][h1], b2), w[4..7] = spatial MLP
                           (W1 [h2][C], b1, w2 [h2], b2 [1]); BatchNorm folded by the caller.
                           i[3] = 1 + id of a buffer (>= 16 channels) that already holds gelu(W1 . s + b1), written by a 1x1 conv
                           op over the spatial-map source (src in mode 0, src2 in mode 1): the op then only applies w2 / b2 and
                           the gated sum — the 180 -> 11 matrix product per pixel runs on the tensor cores; 0: computed here   */
  ,
  RSB_OP_DYSAMPLE = 6   /* DySample head (resselt/utilities/dysample.py:46-83) with its 1x1 end_conv fused, written to the caller's
                           NCHW output: src = features [C] on the low-res grid, src2 = 0.5 * offset(x) * sigmoid(scope(x))
                           [2 * groups * s^2 channels, produced by two 1x1 conv ops], dst_buf = RSB_EXTERNAL_OUTPUT;
                           i[0] groups, i[1] s (up-sampling factor of this head), i[2] out channels, i[3] != 0: src already
                           holds the per-group end_conv projections z[g*4 + o] = sum_{c in group g} W[o][c] x[c] (4 channels
                           per group, produced by a 1x1 conv op; sampling and end_conv commute) and w[1] is ignored;
                           i[4] > 0: src2 holds [0.5 * offset(x) | scope(x)] (scope i[4] channels after offset, one conv op for
                           both) and the op applies the sigmoid gate itself;
                           w[0] = init_pos [2 * groups * s^2], w[1] = end_conv weight [out][C], w[2] = end_conv bias [out].
                           Sampling position of output pixel (h*s+i, w*s+j), group g: (w + off_x, h + off_y) in input pixels,
                           clamped to the image (grid_sample bilinear, align_corners=False, padding_mode='border')            */
  ,
  /* ops of the gated-CNN SPAN descendants (RTMoSR: /root/reference/resselt/archs/rtmosr/arch.py) */
  RSB_OP_RMSNORM = 7,        /* dst = src / (|src|_2 / sqrt(C) + f[0]) * w[0] + w[1] per pixel (channels-first RMSNorm, arch.py:25-38) */
  RSB_OP_UNSHUFFLE_POOL = 8, /* src: C channels (C % 8 == 0) on a grid twice as fine as dst's; dst: [PixelUnshuffle(2)(src) (4C channels,
                                channel c*4 + i*2 + j) | MaxPool2d(2)(src) (C channels)] = 5C channels: the two inputs of
                                ParPixelUnshuffle (arch.py:284-292) in one pass; `channels` = C                              */
  RSB_OP_SE_SHUFFLE = 9      /* src: C channels (C % 32 == 0); dst: C/4 channels on the twice finer grid =
                                PixelShuffle(2)(src * gate), gate = Hardsigmoid(W2 . ReLU(W1 . mean_hw(src) + b1) + b2) per channel
                                (CSELayer, arch.py:7-21) when i[0] = hidden width > 0 (w[0..3] = W1 [hidden][C], b1, W2 [C][hidden],
                                b2), plain PixelShuffle(2) when i[0] == 0; `channels` = C                                       */
  ,
  /* ops of GateRV3's MetaGated block (/root/reference/resselt/archs/gaterv3/arch.py:640-667) */
  RSB_OP_CHAN_GATE = 10,     /* dst = src * ((w[0] . mean_hw(src) + w[1]) * w[2])[c] + src2: simplified channel attention
                                `x * sca(x)` (AdaptiveAvgPool2d(1) -> Conv1x1, w[0] = [C][C], w[1] = bias [C]) times gamma0 (w[2] = [C])
                                plus the block's shortcut (src2, required); C % 8 == 0, C <= 1024                                */
  RSB_OP_CHAN_AFFINE = 11    /* dst = src * w[0][c] (+ src2 when given): `glob(x) * gamma1 + x`; C % 8 == 0                      */
};

typedef struct rsb_op_desc {
  int32_t kind;
  int32_t src_buf, src_ch_off;
  int32_t src2_buf, src2_ch_off; /* RSB_NO_BUFFER when unused */
  int32_t dst_buf, dst_ch_off;
  int32_t channels;
  int32_t i[8];
  float f[4];
  const float* w[8]; /* host arrays, copied */
  int64_t wn[8];     /* their element counts */
} rsb_op_desc;

int rsb_plan_add_op(rsb_plan* plan, const rsb_op_desc* desc);

int rsb_version(void);
/* sizeof the descriptor structs as this library was compiled (0 rsb_conv_desc, 1 rsb_groupnorm_desc, 2 rsb_op_desc, 3 rsb_op_info):
 * lets a foreign-language binding check its own struct layout before the first call */
int rsb_abi_struct_size(int which);
const char* rsb_last_error(void);
/* number of CUDA devices visible to the library (0 on a CPU-only host; never fails) */
int rsb_device_count(void);

/* compute_dtype: RSB_BF16 (tcgen05 tensor-core path, bf16 storage, fp32 accumulate)
 *                RSB_F32  (CUDA-core FFMA path, fp32 storage) */
int rsb_plan_create(int compute_dtype, int in_channels, int out_channels, int upscale, rsb_plan** out);
/* Buffer grids are (H / divisor * scale) x (W / divisor * scale) from then on (call before the first rsb_plan_add_*): lets a plan
 * hold buffers COARSER than its input (RTMoSR's half-resolution branch: divisor 2, full-resolution buffers get scale 2).  The
 * caller's H and W must be multiples of the divisor.  Default 1. */
int rsb_plan_set_base_divisor(rsb_plan* plan, int divisor);
int rsb_plan_destroy(rsb_plan* plan);

/* Activation buffer on the grid (H*scale) x (W*scale); returns its id in *buf_id. */
int rsb_plan_add_buffer(rsb_plan* plan, int channels, int scale, int* buf_id);
int rsb_plan_add_conv(rsb_plan* plan, const rsb_conv_desc* desc);
int rsb_plan_add_groupnorm(rsb_plan* plan, const rsb_groupnorm_desc* desc);

/* Pack the weights (bf16 UMMA layout / fp32) and upload them to `device`. */
int rsb_plan_finalize(rsb_plan* plan, int device);

int rsb_plan_num_ops(const rsb_plan* plan);
/* bf16 plans: convolutions that do not fit the tensor-core kernels and run on the CUDA-core kernel instead (0 for every
 * BASELINE configuration; a warning is printed to stderr once per process when it is not) */
int rsb_plan_num_direct_convs(const rsb_plan* plan);

/* What one op of a finalised plan runs as (in the mode of the last forward, at the shape it bound): for launch accounting and
 * per-kernel timing (bench.py times every launch unit inside a CUDA graph and reports the dominant kernel's roofline). */
enum rsb_kernel_id {
  RSB_K_CONV_DIRECT = 0, /* CUDA-core FFMA conv (the whole fp32 plan)                       */
  RSB_K_CONV_TC = 1,     /* tcgen05 tile kernel (1x1 convs / linears, wide 3x3)             */
  RSB_K_CONV_RS = 2,     /* tcgen05 row-streaming 3x3                                       */
  RSB_K_CONV_LK = 3,     /* tcgen05 row-streaming K x K                                     */
  RSB_K_CONV_PAIR = 4,   /* two 3x3 convs fused, intermediate rows in shared memory         */
  RSB_K_GROUPNORM = 5,
  RSB_K_LAYERNORM = 6,
  RSB_K_DWCONV3 = 7,
  RSB_K_WINATTN = 8,
  RSB_K_CHANATTN = 9,
  RSB_K_AIM = 10,
  RSB_K_DYSAMPLE = 11,
  RSB_K_RMSNORM = 12,
  RSB_K_UNSHUFFLE_POOL = 13,
  RSB_K_SE_SHUFFLE = 14,
  RSB_K_CHAN_GATE = 15,
  RSB_K_CHAN_AFFINE = 16
};
typedef struct rsb_op_info {
  int32_t kind;       /* 0 convolution, 1 GroupNorm, 2 token / attention op                                        */
  int32_t kernel;     /* rsb_kernel_id of the (main) kernel                                                        */
  int32_t fused_next; /* k > 0: this conv and the next k ops run as ONE launch (a fused pair; an N-split group of convs over
                         the same source); the totals below cover all of them and the followers report launches == 0 */
  int32_t launches;   /* kernels launched for this op (0 for the second op of a fused pair)                         */
  double flops;       /* 2 * MACs at the bound shape (convolutions; 0 otherwise)                                    */
  double bytes;       /* algorithmic HBM bytes at the bound shape: inputs + residuals read, outputs written (convs) */
} rsb_op_info;
int rsb_plan_op_info(const rsb_plan* plan, int op_index, rsb_op_info* out);
/* enable != 0: every op of rsb_plan_forward[_ops] runs inside an NVTX range "rsb op <index> <kernel>" (for nsys / ncu --nvtx;
 * off by default: the names are formatted per launch) */
int rsb_plan_set_nvtx(rsb_plan* plan, int enable);
const char* rsb_kernel_name(int kernel_id);
/* kernels launched by one rsb_plan_forward call (for launch accounting) */
int rsb_plan_launches_per_forward(const rsb_plan* plan);
/* 2 * MACs of all convolutions for an n x h x w input */
int rsb_plan_flops(const rsb_plan* plan, int n, int h, int w, double* flops);

int rsb_plan_workspace_bytes(const rsb_plan* plan, int n, int h, int w, size_t* bytes);

/* x: contiguous NCHW [n][in_channels][h][w] of x_dtype on the plan's device;
 * y: contiguous NCHW [n][out_channels][h*upscale][w*upscale] of y_dtype;
 * workspace: 1024-byte aligned device memory of at least rsb_plan_workspace_bytes() bytes.
 * stream: a cudaStream_t passed as void*.  force_direct selects the conv kernels: 0 = fastest available (row-streaming 3x3,
 * tile kernel), 1 = every conv on the CUDA-core kernel (debug cross-check of the tensor-core kernels), 2 = tensor-core tile
 * kernel only (no row-streaming 3x3 kernel; cross-check of the two tensor-core formulations), 3 = row-streaming kernel for
 * every eligible 3x3 conv even where the tile kernel would be preferred (small images), 4 = like 0 but chains of 48-channel
 * 3x3 convs run as fused conv pairs (two convs per launch, the intermediate rows kept in shared memory; bit-identical to 0;
 * measured on B200 it is bound by shared-memory bandwidth at the speed of the two HBM-bound single launches, so it is opt-in),
 * 5 = same as 0. */
int rsb_plan_forward(rsb_plan* plan, const void* x, int x_dtype, int n, int h, int w, void* y, int y_dtype,
                     void* workspace, size_t workspace_bytes, void* stream, int force_direct);

/* Same as rsb_plan_forward but only runs ops [op_begin, op_end) of the plan (ops are numbered in the order they were
 * added; a conv lowered to im2col-pack + 1x1 counts as one op).  For per-layer timing and profiling; the buffers
 * keep whatever the previous calls left in them. */
int rsb_plan_forward_ops(rsb_plan* plan, const void* x, int x_dtype, int n, int h, int w, void* y, int y_dtype,
                         void* workspace, size_t workspace_bytes, void* stream, int force_direct, int op_begin,
                         int op_end);

/* Copy a plan buffer (after a forward) into a dense fp32 NCHW device array, for layer-level tests. */
int rsb_plan_read_buffer(rsb_plan* plan, int buf_id, int ch_off, int channels, float* dst_nchw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RESSELT_B200_H */
