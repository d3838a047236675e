// Host side of the C ABI (include/resselt_b200.h): a "plan" is the layer program of one model —
// activation buffers + fused conv / GroupNorm ops — built once by the Python architecture plugin,
// finalised (weights packed + uploaded) on one device, and replayed on the caller's stream.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"

namespace {

thread_local std::string g_last_error;

// NVTX range around the launches of one op (per-layer ranges for nsys / ncu --nvtx); only when the caller
// switched them on with rsb_plan_set_nvtx — the name is formatted per launch.
struct NvtxRange {
  bool on;
  NvtxRange(bool enable, const char* kernel, int op) : on(enable) {
    if (on) {
      char name[64];
      snprintf(name, sizeof name, "rsb op %d %s", op, kernel);
      nvtxRangePushA(name);
    }
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  return fail((int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
#define RSB_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return fail_cuda(_e, #expr);  \
  } while (0)

constexpr size_t kWsAlign = 1024;
constexpr size_t kMaxSmem = 227 * 1024;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Makes `device` current for a scope and restores the caller's device on every exit path (error returns included).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

uint16_t f32_to_bf16(float f) {  // round-to-nearest-even, same as __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

struct Buffer {
  int channels, planes, scale;
  size_t offset = 0;  // bytes into the workspace (valid after bind)
};

struct ConvOp {
  rsb_conv_desc d;
  std::vector<float> w, b, slopes, border;
  float* d_border = nullptr;  // [16][max(npad, cpad32)]
  float* d_lnsum = nullptr;   // [max(npad, cpad32)] row sums of the rounded weights (LayerNorm fold)
  int scale = 1;  // grid of this conv relative to the input
  int cin_pad16 = 0, npad = 0, cpad32 = 0, cin_planes = 0;
  bool tc_ok = false;
  int stages = 0, kchunk = 0;
  // bf16 plans lower a small-Cin conv on the caller's NCHW tensor to [im2col pack -> 1x1 tensor-core conv]:
  // pack_buf is a hidden planar buffer with pack_k = pad16(cin*kh*kw) channels
  int pack_buf = -1, pack_k = 0;
  // pack_planar: the hidden buffer holds the normalised input itself (channels padded to 16) and the conv keeps its
  // kernel extents, so a 3x3 stem runs on the row-streaming kernel; otherwise the buffer is the im2col matrix
  bool pack_planar = false;
  // geometry the tensor-core kernel sees (differs from d.* only for packed convs)
  int tc_src_buf = -1, tc_src_ch_off = 0, tc_cin = 0, tc_kh = 0, tc_kw = 0;
  rsb::PackParams pk;
  void* d_wtc = nullptr;
  uint32_t wbytes_tc = 0;
  float* d_wdirect = nullptr;
  float* d_bias = nullptr;
  float* d_slopes = nullptr;
  // row-streaming 3x3 kernel (conv_rs.cu): eligibility is decided at finalize, use at bind (needs W % 8 == 0)
  bool rs_elig = false, rs_ready = false, rs_pref = false;
  int rs_stages = 0;
  void* d_wrs = nullptr;
  uint32_t wbytes_rs = 0;
  // large-kernel row-streaming kernel (conv_lk.cu): K x K (odd, 5..17), <= 16 output channels
  bool lk_elig = false, lk_ready = false;
  int lk_stages = 0;
  void* d_wlk = nullptr;
  uint32_t wbytes_lk = 0;
  rsb::ConvLkParams lkp;
  // fused pair kernel (conv_pair.cu): this conv is the head (A) of a pair with the next op; decided at finalize
  bool pair_head = false, pair_ready = false;
  bool pair_store = false;  // a later op reads this conv's output: the pair kernel writes it to its buffer as well
  // N-split group (ConvTcParams::nsplit): this conv heads a run of group_n consecutive convs over the same source that go out as
  // one launch (gtcp); `grouped` marks the others
  int group_n = 1;
  bool grouped = false;
  rsb::ConvTcParams gtcp;
  // bound state
  CUtensorMap map, map_rs, map_res;  // map_res: the pair's second conv's residual rows (L2 prefetch)
  rsb::ConvTcParams tcp;
  rsb::ConvRsParams rsp;
  rsb::ConvPairParams prp;
  rsb::ConvDirectParams dp;
};

struct GnOp {
  rsb_groupnorm_desc d;
  std::vector<float> gamma, beta;
  int scale = 1;
  float* d_gamma = nullptr;
  float* d_beta = nullptr;
  size_t partial_offset = 0;
  int blocks = 0;
  rsb::GroupNormParams gp;
};

struct AuxOp {
  rsb_op_desc d;
  std::vector<float> w[8];
  float* dw[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int scale = 1;
  rsb::TokenOpParams tok;
  rsb::WinAttnParams win;
  rsb::ChanAttnParams chan;
  rsb::AimParams aim;
  rsb::DySampleParams dys;
  rsb::SeParams se;
};

struct Op {
  int kind;  // 0 conv, 1 groupnorm, 2 aux (rsb_op_desc)
  int index;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

struct Cleared {
  void* ws;
  int n, h, w;
};

struct rsb_plan {
  int dtype, in_ch, out_ch, upscale;
  std::vector<Cleared> cleared;  // (workspace, shape) pairs whose padded planes have been zeroed
  std::vector<Buffer> bufs;
  std::vector<ConvOp> convs;
  std::vector<GnOp> gns;
  std::vector<AuxOp> auxs;
  std::vector<Op> ops;
  bool finalized = false;
  int device = -1;
  int num_sms = 0;
  int num_direct = 0;  // bf16 plans: convs that fell back to the CUDA-core kernel
  int info_mode = 0;   // force_direct of the last forward: what rsb_plan_op_info describes
  bool nvtx = false;   // rsb_plan_set_nvtx: one NVTX range per op
  int base_div = 1;    // rsb_plan_set_base_divisor: buffer grids are (H / base_div * scale) x (W / base_div * scale)
  std::mutex mu;
  // binding
  int bn = 0, bh = 0, bw = 0;
  void* bws = nullptr;
  size_t elem() const { return dtype == RSB_BF16 ? 2 : 4; }
};

namespace {

size_t buffer_bytes(const rsb_plan* p, const Buffer& b, int n, int h, int w) {
  return align_up((size_t)n * b.planes * (size_t)(h * b.scale) * (size_t)(w * b.scale) * 8 * p->elem(), kWsAlign);
}

int check_buf(const rsb_plan* p, int id, int ch_off, int channels, const char* what) {
  if (id < 0 || id >= (int)p->bufs.size()) return fail(RSB_ERR_INVALID, "%s: unknown buffer id %d", what, id);
  if (ch_off % 8 != 0) return fail(RSB_ERR_INVALID, "%s: channel offset %d is not a multiple of 8", what, ch_off);
  if (ch_off + channels > p->bufs[id].planes * 8)
    return fail(RSB_ERR_INVALID, "%s: channels [%d, %d) exceed buffer %d (%d channels)", what, ch_off, ch_off + channels, id,
                p->bufs[id].planes * 8);
  return 0;
}

constexpr int kAuxBlocks = 128;  // partial-sum blocks of the global reductions (channel attention Gram, AIM pooling)

size_t aux_scratch_bytes(const AuxOp& a, int n) {
  const rsb_op_desc& d = a.d;
  if (d.kind == RSB_OP_CHANATTN) {
    const int heads = d.i[0], hd = d.channels / heads;
    return ((size_t)n * heads * kAuxBlocks * (hd * hd + 2 * hd) + (size_t)n * heads * hd * hd) * sizeof(float);
  }
  if (d.kind == RSB_OP_AIM) {
    const int cpad = ceil_div(d.channels, 8) * 8;
    return ((size_t)n * kAuxBlocks * cpad + (size_t)n * cpad) * sizeof(float);
  }
  if ((d.kind == RSB_OP_SE_SHUFFLE && d.i[0] > 0) || d.kind == RSB_OP_CHAN_GATE) return ((size_t)n * kAuxBlocks * d.channels + (size_t)n * d.channels) * sizeof(float);
  return 0;
}

int layout(const rsb_plan* p, int n, int h, int w, std::vector<size_t>* offsets, std::vector<size_t>* gn_offsets,
           size_t* total, size_t* aux_offset = nullptr) {
  size_t off = 0;
  if (offsets) offsets->clear();
  for (const Buffer& b : p->bufs) {
    if (offsets) offsets->push_back(off);
    off += buffer_bytes(p, b, n, h, w);
  }
  if (gn_offsets) gn_offsets->clear();
  for (const GnOp& g : p->gns) {
    if (gn_offsets) gn_offsets->push_back(off);
    off += align_up((size_t)n * g.d.groups * 1024 * 2 * sizeof(double), kWsAlign);
  }
  // one scratch region shared by all aux ops (they run one after another on the stream)
  size_t aux = 0;
  for (const AuxOp& a : p->auxs) aux = std::max(aux, aux_scratch_bytes(a, n));
  if (aux_offset) *aux_offset = off;
  off += align_up(aux, kWsAlign);
  *total = off == 0 ? kWsAlign : off;
  return 0;
}

// ---- fused conv pairs (conv_pair.cu)
struct Region {
  int buf, c0, c1;  // channels [c0, c1) of a plan buffer (whole planes)
};
inline bool overlaps(const Region& a, const Region& b) { return a.buf >= 0 && a.buf == b.buf && a.c0 < b.c1 && b.c0 < a.c1; }
inline bool covers(const Region& a, const Region& b) { return a.buf >= 0 && a.buf == b.buf && a.c0 <= b.c0 && a.c1 >= b.c1; }
inline Region conv_src(const ConvOp& c) { return {c.d.src_buf, c.d.src_ch_off, c.d.src_ch_off + c.cin_pad16}; }
inline Region conv_dst(const ConvOp& c) { return {c.d.dst_buf, c.d.dst_ch_off, c.d.dst_ch_off + ceil_div(c.d.cout, 8) * 8}; }

// Is `r` (written by op `writer`) never read by an op after `last_reader`?  Scans forward until an op overwrites all of it.
bool region_dead_after(const rsb_plan* p, const Region& r, size_t last_reader) {
  for (size_t j = last_reader + 1; j < p->ops.size(); ++j) {
    const Op& op = p->ops[j];
    if (op.kind == 0) {
      const ConvOp& c = p->convs[op.index];
      const rsb_conv_desc& d = c.d;
      if (overlaps(conv_src(c), r)) return false;
      if (d.ln_fold && d.ln_stats_buf == r.buf) return false;
      if (d.combine != RSB_COMB_NONE) {
        const int ch = ceil_div(d.cout, 8) * 8;
        if (overlaps({d.res1_buf, d.res1_ch_off, d.res1_ch_off + ch}, r)) return false;
        if (d.res2_buf >= 0 && overlaps({d.res2_buf, d.res2_ch_off, d.res2_ch_off + ch}, r)) return false;
      }
      const bool plain_dst = d.dst_buf >= 0 && d.dst_ps <= 1 && d.dst2_buf < 0;
      if (plain_dst && covers(conv_dst(c), r)) return true;
      if (d.dst_buf == r.buf || d.dst2_buf == r.buf) return false;  // partial overwrite: keep it simple, call it live
    } else if (op.kind == 1) {
      const rsb_groupnorm_desc& d = p->gns[op.index].d;
      if (d.src_buf == r.buf || d.dst_buf == r.buf || d.skip_buf == r.buf) return false;
    } else {
      const rsb_op_desc& d = p->auxs[op.index].d;
      if (d.src_buf == r.buf || d.src2_buf == r.buf || d.dst_buf == r.buf) return false;
      if (d.kind == RSB_OP_AIM && d.i[3] - 1 == r.buf) return false;
    }
  }
  return true;  // the next forward rewrites it before anything reads it
}

// Cut the plan's chains of row-streamable 3x3 convs into fused pairs (conv_pair.cu): conv i becomes the head (A) of a pair
// with conv i + 1 (B) when B reads exactly what A writes.  A may end in an activation or in the SPAB gate; when a later op
// still reads A's output the pair kernel also writes it to its buffer (pair_store), otherwise the map never reaches HBM.
// Greedy from the front: SPAN's conv_1, 6 x (c1_r, c2_r, c3_r), conv_2 become ten pairs.
void find_pairs(rsb_plan* p) {
  static const bool no_pair = rsb::rsb_env("RSB_NO_PAIR") != nullptr;
  if (no_pair || p->dtype != RSB_BF16) return;
  for (size_t i = 0; i + 1 < p->ops.size(); ++i) {
    if (p->ops[i].kind != 0 || p->ops[i + 1].kind != 0) continue;
    ConvOp& a = p->convs[p->ops[i].index];
    ConvOp& b = p->convs[p->ops[i + 1].index];
    const rsb_conv_desc &da = a.d, &db = b.d;
    if (!a.rs_elig || !b.rs_elig || a.tc_cin > 64 || a.npad > 64 || !a.border.empty() || !b.border.empty()) continue;
    if ((a.pack_buf >= 0 && !a.pack_planar) || b.pack_buf >= 0 || a.scale != b.scale) continue;
    if ((da.combine != RSB_COMB_NONE && da.combine != RSB_COMB_SPAB_GATE) || da.res2_buf >= 0 || da.act == RSB_ACT_PRELU) continue;
    if (da.dst_buf < 0 || da.dst_ps > 1 || da.dst2_buf >= 0) continue;
    if (da.cout != a.npad || b.npad != a.npad) continue;  // one UMMA N for both convs, no padded channels in the ring
    if (db.src_buf != da.dst_buf || db.src_ch_off != da.dst_ch_off || db.cin != da.cout || db.src_upsample2) continue;
    if (db.dst_buf < 0 || db.dst_ps > 1 || db.dst2_buf >= 0 || db.res2_buf >= 0 || db.act == RSB_ACT_PRELU) continue;
    const Region mid = conv_dst(a), bdst = conv_dst(b), asrc = conv_src(a);
    if (overlaps(bdst, asrc) || overlaps(bdst, mid)) continue;  // B's rows are written while A still reads its source
    if (da.combine != RSB_COMB_NONE) {
      const Region res = {da.res1_buf, da.res1_ch_off, da.res1_ch_off + ceil_div(da.cout, 8) * 8};
      if (overlaps(res, mid) || overlaps(res, bdst)) continue;
    }
    if (db.combine != RSB_COMB_NONE) {
      const Region res = {db.res1_buf, db.res1_ch_off, db.res1_ch_off + ceil_div(db.cout, 8) * 8};
      if (overlaps(res, mid) || overlaps(res, bdst)) continue;
    }
    const bool store = !region_dead_after(p, mid, i + 1);
    if (!rsb::conv_pair_supported(a.tc_cin, a.npad, da.act, da.combine, store, db.act, db.combine)) continue;
    if (rsb::conv_pair_smem_bytes(a.tc_cin, a.npad) > kMaxSmem) continue;
    a.pair_head = true;
    a.pair_store = store;
    ++i;  // pairs do not overlap
  }
}

void fill_epi(rsb_plan* p, ConvOp& c, int n, int H, int W, uint8_t* ws, rsb::Epi& e) {
  const rsb_conv_desc& d = c.d;
  memset(&e, 0, sizeof e);
  e.bias = c.d_bias;
  e.slopes = c.d_slopes;
  e.border_bias = c.d_border;
  e.bb_stride = std::max(c.npad, c.cpad32);
  if (d.ln_fold) {
    e.ln_stats = ws + p->bufs[d.ln_stats_buf].offset;
    e.ln_stride = p->dtype == RSB_BF16 ? 16 : 32;  // one 8-channel pixel chunk
    e.ln_planes = p->bufs[d.ln_stats_buf].planes;
    e.ln_rowsum = c.d_lnsum;
    e.ln_raw = d.ln_fold == 2 ? 1 : 0;
    e.ln_inv_n = 1.0f / (float)d.cin, e.ln_eps = d.ln_eps;
  }
  if (d.ln_out) {
    e.ln_out = ws + p->bufs[d.ln_out_buf].offset;
    e.ln_out_stride = 16;  // bf16 plans only: one 8-channel pixel chunk
    e.ln_out_planes = p->bufs[d.ln_out_buf].planes;
  }
  e.act = d.act;
  e.act_param = d.act_param;
  e.combine = d.combine;
  e.alpha = d.alpha, e.beta1 = d.beta1, e.beta2 = d.beta2;
  if (d.combine != RSB_COMB_NONE) {
    const Buffer& r = p->bufs[d.res1_buf];
    e.res1 = ws + r.offset;
    e.res1_planes = r.planes;
    e.res1_plane0 = d.res1_ch_off / 8;
    if (d.combine == RSB_COMB_AXPY && d.res2_buf >= 0) {
      const Buffer& r2 = p->bufs[d.res2_buf];
      e.res2 = ws + r2.offset;
      e.res2_planes = r2.planes;
      e.res2_plane0 = d.res2_ch_off / 8;
    }
  }
  e.cout = d.cout;
  e.H = H, e.W = W;
  if (d.dst_buf == RSB_EXTERNAL_OUTPUT) {
    e.dst_external = 1;
    e.ps = d.ps;
    e.out_ch = d.cout / (d.ps * d.ps);
    e.add_base = d.add_base;
    e.base_ch = p->in_ch;
    e.out_scale = d.out_scale;
    for (int i = 0; i < 4; ++i) e.out_mean[i] = d.out_mean[i];
  } else {
    const Buffer& b = p->bufs[d.dst_buf];
    e.dst = ws + b.offset;
    e.dst_planes = b.planes;
    e.dst_plane0 = d.dst_ch_off / 8;
    e.dst_ps = d.dst_ps > 1 ? d.dst_ps : 1;
    e.phase_ch = (e.dst_ps > 1 && d.dst_phase < 0) ? d.cout / (e.dst_ps * e.dst_ps) : d.cout;
    e.phase0 = (e.dst_ps > 1 && d.dst_phase >= 0) ? d.dst_phase : 0;
    if (e.dst_ps == 1 && d.dst2_buf >= 0) {
      const Buffer& b2 = p->bufs[d.dst2_buf];
      e.dst2 = ws + b2.offset;
      e.dst2_planes = b2.planes;
      e.dst2_plane0 = d.dst2_ch_off / 8;
      e.split_ch = d.split_ch;
    }
    e.simple = (e.dst_ps == 1 && e.dst2 == nullptr && e.res2 == nullptr && e.border_bias == nullptr && e.ln_stats == nullptr) ? 1 : 0;
  }
}

int bind(rsb_plan* p, int n, int h, int w, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  std::vector<size_t> offs, gn_offs;
  size_t total, aux_off = 0;
  layout(p, n, h, w, &offs, &gn_offs, &total, &aux_off);
  if (ws_bytes < total) return fail(RSB_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, total);
  if ((uintptr_t)workspace % kWsAlign != 0) return fail(RSB_ERR_WORKSPACE, "workspace must be %zu-byte aligned", kWsAlign);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  for (size_t i = 0; i < p->bufs.size(); ++i) p->bufs[i].offset = offs[i];
  // padded planes must hold finite values (they meet zero weights): clear once per (workspace, shape).  A caller that
  // alternates between a few shapes with one workspace each (edge tiles of a tiled forward) re-binds without re-clearing:
  // the layout for a shape is fixed, and the ops never write anything but finite values into padded channels.
  {
    const Cleared key = {workspace, n, h, w};
    bool seen = false;
    for (const Cleared& c : p->cleared) seen = seen || (c.ws == key.ws && c.n == n && c.h == h && c.w == w);
    if (!seen) {
      RSB_CUDA(cudaMemsetAsync(workspace, 0, total, stream));
      // a different shape on the same memory invalidates what was recorded for it
      p->cleared.erase(std::remove_if(p->cleared.begin(), p->cleared.end(), [&](const Cleared& c) { return c.ws == key.ws; }), p->cleared.end());
      if (p->cleared.size() >= 8) p->cleared.erase(p->cleared.begin());
      p->cleared.push_back(key);
    }
  }

  for (ConvOp& c : p->convs) {
    const rsb_conv_desc& d = c.d;
    const int H = h * c.scale, W = w * c.scale;
    if (c.tc_ok) {
      EncodeTiledFn enc = get_encode_fn();
      if (!enc) return fail(RSB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled entry point not available");
      const Buffer& sb = p->bufs[c.tc_src_buf];
      const int HT = rsb::kTileH + c.tc_kh - 1, WT = rsb::kTileW + c.tc_kw - 1;
      // 1x1 convs between planar buffers have no neighbourhood: view the image as (H*W/8) rows of 8 pixels, so that a
      // 16 x 8 tile is 128 CONSECUTIVE pixels — 2 KB contiguous per plane for the TMA load, the residual read and the store
      // instead of sixteen 128-byte pieces one image row apart.  Same linear addresses, same arithmetic, different tile shape.
      static const bool no_linear = rsb::rsb_env("RSB_NO_LINEAR1X1") != nullptr;
      const bool linear = !no_linear && c.tc_kh == 1 && c.tc_kw == 1 && d.dst_buf >= 0 && d.dst_ps <= 1 && ((long long)H * W) % 8 == 0;
      const int Hm = linear ? (int)((long long)H * W / 8) : H, Wm = linear ? 8 : W;
      cuuint64_t dims[4] = {(cuuint64_t)Wm * 8, (cuuint64_t)Hm, (cuuint64_t)sb.planes, (cuuint64_t)n};
      cuuint64_t strides[3] = {(cuuint64_t)Wm * 16, (cuuint64_t)Wm * 16 * Hm, (cuuint64_t)Wm * 16 * Hm * sb.planes};
      cuuint32_t box[4] = {(cuuint32_t)(8 * WT), (cuuint32_t)HT, (cuuint32_t)(c.kchunk / 8), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&c.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ws + sb.offset, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(RSB_ERR_INVALID, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
      rsb::ConvTcParams& t = c.tcp;
      memset(&t, 0, sizeof t);
      t.n = n, t.H = Hm, t.W = Wm;
      t.tiles_x = ceil_div(Wm, rsb::kTileW), t.tiles_y = ceil_div(Hm, rsb::kTileH);
      t.num_tiles = t.tiles_x * t.tiles_y * n;
      t.cin = c.tc_cin, t.npad = c.npad;
      t.kchunk = c.kchunk, t.nchunks = c.tc_cin / c.kchunk;
      {
        // K-chunked tiles: both issuing warps share ONE ring of stages in chunk order.  The seen[] counters of the kernel make
        // every full-barrier wait phase-exact, so two warps are safe for any chunk / stage count (measured +2..4 % on RealPLKSR,
        // SwinIR, DAT over a single issuing warp).  RSB_TC_SOLO=1 restores one issuing warp whenever 2 * chunks > stages.
        static const bool solo_env = rsb::rsb_env("RSB_TC_SOLO") != nullptr;
        t.solo_issue = solo_env && t.nchunks > 1 && 2 * t.nchunks > c.stages;
      }
      t.kh = c.tc_kh, t.kw = c.tc_kw;
      const bool im2col = c.pack_buf >= 0 && !c.pack_planar;
      t.pad_t = im2col ? 0 : d.pad_t, t.pad_l = im2col ? 0 : d.pad_l;
      t.src_plane0 = c.tc_src_ch_off / 8;
      t.wpack = c.d_wtc, t.wbytes = c.wbytes_tc;
      t.stages = c.stages;
      t.stage_bytes = (uint32_t)HT * WT * c.kchunk * 2u;
      t.num_acc = rsb::conv_tc_num_acc(c.npad);
      t.acc_stride = (uint32_t)c.npad;
      uint32_t cols = 32;
      while (cols < (uint32_t)t.num_acc * c.npad) cols <<= 1;
      t.tmem_cols = cols;
      fill_epi(p, c, n, Hm, Wm, ws, t.epi);
      // each CTA streams a contiguous run of rows and pays ~2 halo rows per run: worth it from ~8 rows per CTA
      c.rs_ready = c.rs_elig && W % 8 == 0;
      c.lk_ready = c.lk_elig && W % 8 == 0;
      c.rs_pref = (long long)n * ceil_div(W, 128) * H >= 8ll * p->num_sms;
      if (c.rs_ready || c.lk_ready) {
        // the same tensor viewed as [n][plane][H][W/8][8 px x 8 ch]: a box of 18 pixel groups of one row lands as
        // [plane][18][128 B], i.e. 144 consecutive pixels per plane
        cuuint64_t dims5[5] = {64, (cuuint64_t)(W / 8), (cuuint64_t)H, (cuuint64_t)sb.planes, (cuuint64_t)n};
        cuuint64_t strides5[4] = {128, (cuuint64_t)W * 16, (cuuint64_t)W * 16 * H, (cuuint64_t)W * 16 * H * sb.planes};
        cuuint32_t box5[5] = {64, 18, 1, (cuuint32_t)(c.tc_cin / 8), 1};
        cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
        CUresult r5 = enc(&c.map_rs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ws + sb.offset, dims5, strides5, box5, estr5,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r5 != CUDA_SUCCESS) return fail(RSB_ERR_INVALID, "cuTensorMapEncodeTiled (5-D) failed with CUresult %d", (int)r5);
        rsb::ConvRsParams& q = c.rsp;
        memset(&q, 0, sizeof q);
        q.n = n, q.H = H, q.W = W;
        q.cols = ceil_div(W, 128);
        q.units = n * q.cols * H;
        q.cin = c.tc_cin, q.np = c.npad, q.nslots = std::min(32, 512 / c.npad);
        q.src_plane0 = c.tc_src_ch_off / 8;
        q.wpack = c.d_wrs, q.wbytes = c.wbytes_rs;
        q.stages = c.rs_stages;
        q.stage_bytes = rsb::conv_rs_stage_bytes(c.tc_cin);
        q.epi = t.epi;
        if (c.lk_ready) {
          rsb::ConvLkParams& k = c.lkp;
          memset(&k, 0, sizeof k);
          k.n = n, k.H = H, k.W = W;
          k.cols = ceil_div(W, 128);
          k.units = n * k.cols * H;
          k.cin = c.tc_cin, k.k = d.kh;
          k.src_plane0 = c.tc_src_ch_off / 8;
          k.wpack = c.d_wlk, k.wbytes = c.wbytes_lk;
          k.stages = c.lk_stages;
          k.stage_bytes = rsb::conv_rs_stage_bytes(c.tc_cin);
          k.epi = t.epi;
        }
      }
      if (c.pack_buf >= 0) {
        rsb::PackParams& k = c.pk;
        memset(&k, 0, sizeof k);
        k.n = n, k.H = H, k.W = W, k.cin = d.cin;
        k.kh = c.pack_planar ? 1 : d.kh, k.kw = c.pack_planar ? 1 : d.kw;
        k.pad_t = c.pack_planar ? 0 : d.pad_t, k.pad_l = c.pack_planar ? 0 : d.pad_l;
        k.kplanes = c.pack_k / 8;
        for (int i = 0; i < 4; ++i) k.in_mean[i] = d.in_mean[i];
        k.in_scale = d.in_scale;
        k.dst = ws + sb.offset;
      }
    }
    rsb::ConvDirectParams& q = c.dp;
    memset(&q, 0, sizeof q);
    q.n = n, q.H = H, q.W = W;
    q.cin = d.cin, q.cin_planes = c.cin_planes;
    q.cout = d.cout, q.cpad = c.cpad32;
    q.kh = d.kh, q.kw = d.kw, q.pad_t = d.pad_t, q.pad_l = d.pad_l;
    if (d.src_buf == RSB_EXTERNAL_INPUT) {
      q.src_external = 1;
      for (int i = 0; i < 4; ++i) q.in_mean[i] = d.in_mean[i];
      q.in_scale = d.in_scale;
    } else {
      const Buffer& sb = p->bufs[d.src_buf];
      q.src = ws + sb.offset;
      q.src_planes = sb.planes;
      q.src_plane0 = d.src_ch_off / 8;
      q.src_upsample2 = d.src_upsample2;
    }
    q.wpack = c.d_wdirect;
    fill_epi(p, c, n, H, W, ws, q.epi);
  }
  // ---- N-split groups: runs of up to three consecutive tile-kernel convs that read the same source with the same geometry and
  // differ only in weights / bias / destination channel offset (q, k, v; the halves of an MLP's first linear)
  for (ConvOp& c : p->convs) c.group_n = 1, c.grouped = false;
  {
    static const bool no_split = rsb::rsb_env("RSB_NO_NSPLIT") != nullptr;
    auto tile_only = [&](const ConvOp& c) {
      return c.tc_ok && c.pack_buf < 0 && !(c.rs_ready && c.rs_pref) && !(c.lk_ready && c.rs_pref) && !c.pair_head;
    };
    auto same = [&](const ConvOp& a, const ConvOp& b) {
      const rsb_conv_desc &x = a.d, &y = b.d;
      const rsb::ConvTcParams &s = a.tcp, &t = b.tcp;
      return a.tc_src_buf == b.tc_src_buf && a.tc_src_ch_off == b.tc_src_ch_off && a.tc_cin == b.tc_cin && a.scale == b.scale &&
             x.kh == y.kh && x.kw == y.kw && x.pad_t == y.pad_t && x.pad_l == y.pad_l && x.dst_buf == y.dst_buf && x.dst_buf >= 0 &&
             x.dst_ps <= 1 && y.dst_ps <= 1 && x.dst2_buf < 0 && y.dst2_buf < 0 && x.act == y.act && x.act_param == y.act_param &&
             x.combine == RSB_COMB_NONE && y.combine == RSB_COMB_NONE && a.border.empty() && b.border.empty() &&
             x.ln_fold == y.ln_fold && (!x.ln_fold || (x.ln_stats_buf == y.ln_stats_buf && x.ln_eps == y.ln_eps)) && !x.ln_out && !y.ln_out &&
             a.npad == b.npad && s.stages == t.stages &&
             s.kchunk == t.kchunk && s.solo_issue == t.solo_issue && s.num_acc == t.num_acc && s.acc_stride == t.acc_stride &&
             s.tmem_cols == t.tmem_cols && s.wbytes == t.wbytes && s.stage_bytes == t.stage_bytes && y.dst_ch_off >= x.dst_ch_off;
    };
    for (size_t i = 0; i + 1 < p->ops.size() && !no_split; ++i) {
      if (p->ops[i].kind != 0) continue;
      ConvOp& a = p->convs[p->ops[i].index];
      if (!tile_only(a) || a.d.dst_buf == a.tc_src_buf) continue;
      int k = 1;
      while (k < 3 && i + k < p->ops.size() && p->ops[i + k].kind == 0 && p->num_sms / (k + 1) >= 1) {
        const ConvOp& b = p->convs[p->ops[i + k].index];
        if (!tile_only(b) || !same(a, b)) break;
        // destinations must not overlap one another
        bool clash = false;
        for (int j = 0; j < k && !clash; ++j) clash = overlaps(conv_dst(p->convs[p->ops[i + j].index]), conv_dst(b));
        if (clash) break;
        ++k;
      }
      if (k < 2) continue;
      a.group_n = k;
      a.gtcp = a.tcp;
      a.gtcp.nsplit = k;
      for (int j = 0; j < k; ++j) {
        ConvOp& b = p->convs[p->ops[i + j].index];
        if (j > 0) b.grouped = true;
        rsb::ConvTcParams::Slice& sl = a.gtcp.slice[j];
        sl.wpack = b.d_wtc;
        sl.bias = b.d_bias;
        sl.aux = b.d.ln_fold ? b.d_lnsum : b.d_slopes;
        sl.dst_plane_off = (b.d.dst_ch_off - a.d.dst_ch_off) / 8;
        sl.cout = b.d.cout;
      }
      i += k - 1;
    }
  }
  for (size_t i = 0; i + 1 < p->ops.size(); ++i) {
    if (p->ops[i].kind != 0) continue;
    ConvOp& a = p->convs[p->ops[i].index];
    a.pair_ready = false;
    if (!a.pair_head) continue;
    ConvOp& b = p->convs[p->ops[i + 1].index];
    const int H = h * a.scale, W = w * a.scale;
    const int cols = rsb::conv_pair_cols(W);
    // same eligibility as the row-streaming kernel (W % 8 == 0, enough rows per CTA to amortise the run ends)
    if (!a.rs_ready || !b.rs_ready || (long long)n * cols * H < 8ll * p->num_sms) continue;
    rsb::ConvPairParams& q = a.prp;
    memset(&q, 0, sizeof q);
    q.n = n, q.H = H, q.W = W;
    q.cols = cols, q.units = n * cols * H;
    q.cin0 = a.tc_cin, q.np = a.npad;
    q.src_plane0 = a.tc_src_ch_off / 8;
    q.wpackA = a.d_wrs, q.wbytesA = a.wbytes_rs;
    q.wpackB = b.d_wrs, q.wbytesB = b.wbytes_rs;
    q.stage_bytes = rsb::conv_rs_stage_bytes(a.tc_cin);
    q.biasA = a.d_bias, q.actA = a.d.act, q.actA_param = a.d.act_param;
    q.combA = a.d.combine;
    if (a.d.combine != RSB_COMB_NONE) {
      const Buffer& rb = p->bufs[a.d.res1_buf];
      q.resA = ws + rb.offset, q.resA_planes = rb.planes, q.resA_plane0 = a.d.res1_ch_off / 8;
    }
    if (a.pair_store) {
      const Buffer& db = p->bufs[a.d.dst_buf];
      q.dstA = ws + db.offset, q.dstA_planes = db.planes, q.dstA_plane0 = a.d.dst_ch_off / 8;
    }
    q.epi = b.tcp.epi;
    memset(&a.map_res, 0, sizeof a.map_res);
    const int res_of = b.d.combine != RSB_COMB_NONE ? 1 : (a.d.combine != RSB_COMB_NONE ? 2 : 0);
    if (res_of != 0) {
      const rsb_conv_desc& rd = res_of == 1 ? b.d : a.d;
      EncodeTiledFn enc = get_encode_fn();
      const Buffer& rb = p->bufs[rd.res1_buf];
      cuuint64_t dims5[5] = {64, (cuuint64_t)(W / 8), (cuuint64_t)H, (cuuint64_t)rb.planes, (cuuint64_t)n};
      cuuint64_t strides5[4] = {128, (cuuint64_t)W * 16, (cuuint64_t)W * 16 * H, (cuuint64_t)W * 16 * H * rb.planes};
      cuuint32_t box5[5] = {64, 16, 1, (cuuint32_t)(a.npad / 8), 1};
      cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
      CUresult r5 = enc(&a.map_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ws + rb.offset, dims5, strides5, box5, estr5,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r5 != CUDA_SUCCESS) return fail(RSB_ERR_INVALID, "cuTensorMapEncodeTiled (pair residual) failed with CUresult %d", (int)r5);
      q.res_prefetch = res_of;
      q.res_plane0 = rd.res1_ch_off / 8;
    }
    a.pair_ready = true;
  }
  for (size_t i = 0; i < p->gns.size(); ++i) {
    GnOp& g = p->gns[i];
    const int H = h * g.scale, W = w * g.scale;
    rsb::GroupNormParams& q = g.gp;
    memset(&q, 0, sizeof q);
    q.n = n, q.H = H, q.W = W;
    q.channels = g.d.channels, q.groups = g.d.groups, q.eps = g.d.eps;
    const Buffer& sb = p->bufs[g.d.src_buf];
    q.src = ws + sb.offset, q.src_planes = sb.planes, q.src_plane0 = g.d.src_ch_off / 8;
    const Buffer& db = p->bufs[g.d.dst_buf];
    q.dst = ws + db.offset, q.dst_planes = db.planes, q.dst_plane0 = g.d.dst_ch_off / 8;
    if (g.d.skip_buf >= 0) {
      const Buffer& kb = p->bufs[g.d.skip_buf];
      q.skip = ws + kb.offset, q.skip_planes = kb.planes, q.skip_plane0 = g.d.skip_ch_off / 8;
    }
    q.gamma = g.d_gamma, q.beta = g.d_beta;
    q.partial = reinterpret_cast<double*>(ws + gn_offs[i]);
    const size_t chunks = (size_t)H * W * (g.d.channels / g.d.groups / 8);
    q.blocks_per_group = (int)std::min<size_t>(1024, std::max<size_t>(1, chunks / 2048));
  }
  for (AuxOp& a : p->auxs) {
    const rsb_op_desc& d = a.d;
    // a.scale is the source buffer's grid; the pixel-unshuffle op iterates over its (half as fine) destination grid
    const int gs = d.kind == RSB_OP_UNSHUFFLE_POOL ? a.scale / 2 : a.scale;
    const int H = h * gs, W = w * gs;
    const Buffer& sb = p->bufs[d.src_buf];
    if (d.kind == RSB_OP_DYSAMPLE) {
      rsb::DySampleParams& t = a.dys;
      memset(&t, 0, sizeof t);
      const Buffer& ob = p->bufs[d.src2_buf];
      t.n = n, t.H = H, t.W = W, t.channels = d.channels, t.groups = d.i[0], t.s = d.i[1], t.out_ch = d.i[2];
      t.projected = d.i[3] != 0;
      t.gate_off = d.i[4];
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_plane0 = d.src_ch_off / 8;
      t.off = ws + ob.offset, t.off_planes = ob.planes, t.off_plane0 = d.src2_ch_off / 8;
      t.init_pos = a.dw[0], t.weight = a.dw[1], t.bias = a.dw[2];
      continue;
    }
    const Buffer& db = p->bufs[d.dst_buf];
    uint8_t* scratch = ws + aux_off;
    if (d.kind == RSB_OP_CHAN_GATE) {
      rsb::SeParams& t = a.se;
      memset(&t, 0, sizeof t);
      const Buffer& rb = p->bufs[d.src2_buf];
      t.n = n, t.H = H, t.W = W, t.channels = d.channels;
      t.blocks = (int)std::min<size_t>(kAuxBlocks, std::max<size_t>(1, ((size_t)H * W + 2047) / 2048));
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_plane0 = d.src_ch_off / 8;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_plane0 = d.dst_ch_off / 8;
      t.res = ws + rb.offset, t.res_planes = rb.planes, t.res_plane0 = d.src2_ch_off / 8;
      t.w1 = a.dw[0], t.b1 = a.dw[1], t.w2 = a.dw[2];
      t.partial = reinterpret_cast<float*>(scratch);
      t.gate = t.partial + (size_t)n * kAuxBlocks * d.channels;
      continue;
    }
    if (d.kind == RSB_OP_SE_SHUFFLE) {
      rsb::SeParams& t = a.se;
      memset(&t, 0, sizeof t);
      t.n = n, t.H = H, t.W = W, t.channels = d.channels, t.hidden = d.i[0];
      t.blocks = (int)std::min<size_t>(32, std::max<size_t>(1, ((size_t)H * W + 4095) / 4096));
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_plane0 = d.src_ch_off / 8;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_plane0 = d.dst_ch_off / 8;
      if (t.hidden > 0) {
        t.w1 = a.dw[0], t.b1 = a.dw[1], t.w2 = a.dw[2], t.b2 = a.dw[3];
        t.partial = reinterpret_cast<float*>(scratch);
        t.gate = t.partial + (size_t)n * kAuxBlocks * d.channels;
      }
      continue;
    }
    if (d.kind == RSB_OP_LAYERNORM || d.kind == RSB_OP_DWCONV3 || d.kind == RSB_OP_RMSNORM || d.kind == RSB_OP_UNSHUFFLE_POOL || d.kind == RSB_OP_CHAN_AFFINE) {
      rsb::TokenOpParams& t = a.tok;
      memset(&t, 0, sizeof t);
      t.n = n, t.H = H, t.W = W, t.channels = d.channels;
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_plane0 = d.src_ch_off / 8;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_plane0 = d.dst_ch_off / 8;
      if (d.src2_buf >= 0) {
        const Buffer& s2 = p->bufs[d.src2_buf];
        t.src2 = ws + s2.offset, t.src2_planes = s2.planes, t.src2_plane0 = d.src2_ch_off / 8;
      }
      t.w0 = a.dw[0], t.w1 = a.dw[1], t.f0 = d.f[0], t.i0 = d.i[0];
      t.dense5 = (d.kind == RSB_OP_DWCONV3 && d.i[1] == 5) ? 1 : 0;
    } else if (d.kind == RSB_OP_WINATTN) {
      rsb::WinAttnParams& t = a.win;
      memset(&t, 0, sizeof t);
      const int m = std::max(d.i[1], d.i[2]);
      t.n = n, t.H = H, t.W = W, t.Hp = ceil_div(H, m) * m, t.Wp = ceil_div(W, m) * m;
      t.dim = d.channels, t.heads = d.i[0], t.head_dim = d.channels / d.i[0];
      t.split_h = d.i[1], t.split_w = d.i[2], t.shifted = d.i[3], t.scale = d.f[0];
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_ch_off = d.src_ch_off;
      t.qkv_stride = d.i[4] > 0 ? d.i[4] : d.channels;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_ch_off = d.dst_ch_off;
      t.table0 = a.dw[0], t.table1 = a.dw[1];
      t.head_pad = d.i[5];
    } else if (d.kind == RSB_OP_CHANATTN) {
      rsb::ChanAttnParams& t = a.chan;
      memset(&t, 0, sizeof t);
      t.n = n, t.H = H, t.W = W, t.dim = d.channels, t.heads = d.i[0], t.head_dim = d.channels / d.i[0];
      t.blocks = (int)std::min<size_t>(kAuxBlocks, std::max<size_t>(1, ((size_t)H * W + 255) / 256));
      t.src = ws + sb.offset, t.src_planes = sb.planes, t.src_ch_off = d.src_ch_off;
      t.qkv_stride = d.i[1] > 0 ? d.i[1] : d.channels;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_ch_off = d.dst_ch_off;
      t.temperature = a.dw[0];
      t.partial = reinterpret_cast<float*>(scratch);
      t.attn = t.partial + (size_t)n * t.heads * kAuxBlocks * (t.head_dim * t.head_dim + 2 * t.head_dim);
    } else if (d.kind == RSB_OP_AIM) {
      rsb::AimParams& t = a.aim;
      memset(&t, 0, sizeof t);
      const Buffer& cb = p->bufs[d.src2_buf];
      t.n = n, t.H = H, t.W = W, t.channels = d.channels, t.cpad = ceil_div(d.channels, 8) * 8;
      t.mode = d.i[0], t.ci_hidden = d.i[1], t.si_hidden = d.i[2];
      t.blocks = (int)std::min<size_t>(kAuxBlocks, std::max<size_t>(1, ((size_t)H * W + 1023) / 1024));
      t.att = ws + sb.offset, t.att_planes = sb.planes, t.att_plane0 = d.src_ch_off / 8;
      t.convx = ws + cb.offset, t.convx_planes = cb.planes, t.convx_plane0 = d.src2_ch_off / 8;
      if (t.mode == 0)
        t.pool_src = t.convx, t.pool_planes = t.convx_planes, t.pool_plane0 = t.convx_plane0;
      else
        t.pool_src = t.att, t.pool_planes = t.att_planes, t.pool_plane0 = t.att_plane0;
      t.dst = ws + db.offset, t.dst_planes = db.planes, t.dst_plane0 = d.dst_ch_off / 8;
      t.ci_w1 = a.dw[0], t.ci_b1 = a.dw[1], t.ci_w2 = a.dw[2], t.ci_b2 = a.dw[3];
      t.si_w1 = a.dw[4], t.si_b1 = a.dw[5], t.si_w2 = a.dw[6], t.si_b2 = a.w[7].empty() ? 0.0f : a.w[7][0];
      t.partial = reinterpret_cast<float*>(scratch);
      t.cmap = t.partial + (size_t)n * kAuxBlocks * t.cpad;
      if (d.i[3] > 0) t.hid = ws + p->bufs[d.i[3] - 1].offset, t.hid_planes = p->bufs[d.i[3] - 1].planes;
    }
  }
  p->bn = n, p->bh = h, p->bw = w, p->bws = workspace;
  return 0;
}

}  // namespace

extern "C" {

int rsb_version(void) { return RSB_VERSION; }
int rsb_abi_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(rsb_conv_desc);
    case 1: return (int)sizeof(rsb_groupnorm_desc);
    case 2: return (int)sizeof(rsb_op_desc);
    case 3: return (int)sizeof(rsb_op_info);
    default: return -1;
  }
}
const char* rsb_last_error(void) { return g_last_error.c_str(); }

int rsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int rsb_plan_create(int compute_dtype, int in_channels, int out_channels, int upscale, rsb_plan** out) {
  if (!out) return fail(RSB_ERR_INVALID, "rsb_plan_create: out is NULL");
  if (compute_dtype != RSB_F32 && compute_dtype != RSB_BF16)
    return fail(RSB_ERR_INVALID, "rsb_plan_create: compute dtype must be RSB_F32 or RSB_BF16");
  if (in_channels < 1 || out_channels < 1 || upscale < 1) return fail(RSB_ERR_INVALID, "rsb_plan_create: bad channel/upscale");
  rsb_plan* p = new rsb_plan();
  p->dtype = compute_dtype, p->in_ch = in_channels, p->out_ch = out_channels, p->upscale = upscale;
  *out = p;
  return 0;
}

int rsb_plan_destroy(rsb_plan* p) {
  if (!p) return 0;
  if (p->device >= 0) {
    // also after a finalize that failed half-way: whatever was allocated up to that point is released (null pointers are fine)
    DeviceGuard guard(p->device);
    for (ConvOp& c : p->convs) {
      cudaFree(c.d_wtc), cudaFree(c.d_wrs), cudaFree(c.d_wlk), cudaFree(c.d_wdirect), cudaFree(c.d_bias), cudaFree(c.d_slopes), cudaFree(c.d_border);
      cudaFree(c.d_lnsum);
    }
    for (GnOp& g : p->gns) cudaFree(g.d_gamma), cudaFree(g.d_beta);
    for (AuxOp& a : p->auxs)
      for (int k = 0; k < 8; ++k) cudaFree(a.dw[k]);
    cudaGetLastError();
  }
  delete p;
  return 0;
}

int rsb_plan_set_base_divisor(rsb_plan* p, int divisor) {
  if (!p) return fail(RSB_ERR_INVALID, "rsb_plan_set_base_divisor: NULL plan");
  if (p->finalized || !p->bufs.empty() || !p->ops.empty()) return fail(RSB_ERR_STATE, "rsb_plan_set_base_divisor: call before adding buffers / ops");
  if (divisor < 1 || divisor > 64) return fail(RSB_ERR_INVALID, "rsb_plan_set_base_divisor: divisor must be in [1, 64]");
  p->base_div = divisor;
  return 0;
}

int rsb_plan_add_buffer(rsb_plan* p, int channels, int scale, int* buf_id) {
  if (!p || !buf_id) return fail(RSB_ERR_INVALID, "rsb_plan_add_buffer: NULL argument");
  if (p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_add_buffer: plan already finalized");
  if (channels < 1 || scale < 1) return fail(RSB_ERR_INVALID, "rsb_plan_add_buffer: bad channels/scale");
  Buffer b;
  b.channels = channels;
  b.planes = ceil_div(channels, 16) * 2;  // whole 16-channel K steps
  b.scale = scale;
  p->bufs.push_back(b);
  *buf_id = (int)p->bufs.size() - 1;
  return 0;
}

int rsb_plan_add_conv(rsb_plan* p, const rsb_conv_desc* desc) {
  if (!p || !desc) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: NULL argument");
  if (p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_add_conv: plan already finalized");
  const rsb_conv_desc& d = *desc;
  const bool default_pad = d.pad_t < 0 || d.pad_l < 0;
  if (d.cin < 1 || d.cout < 1 || d.kh < 1 || d.kw < 1 || (default_pad && (d.kh % 2 == 0 || d.kw % 2 == 0)) ||
      (!default_pad && (d.pad_t >= d.kh || d.pad_l >= d.kw)))
    return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: bad cin/cout/kernel (%d,%d,%dx%d)", d.cin, d.cout, d.kh, d.kw);
  if (!d.weight) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: weight is NULL");
  if (d.act < RSB_ACT_NONE || d.act > RSB_ACT_GELU) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: unknown activation %d", d.act);
  if (d.act == RSB_ACT_PRELU && !d.act_slopes) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: PReLU needs act_slopes");
  if (d.combine < RSB_COMB_NONE || d.combine > RSB_COMB_AXPY) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: unknown combine %d", d.combine);
  ConvOp c;
  c.d = d;
  if (default_pad) c.d.pad_t = d.kh / 2, c.d.pad_l = d.kw / 2;
  int scale = p->base_div;  // the caller's input lives on the full grid = base_div x the plan's base grid
  if (d.src_buf == RSB_EXTERNAL_INPUT) {
    if (d.cin != p->in_ch) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: external input has %d channels, conv wants %d", p->in_ch, d.cin);
    if (d.src_upsample2) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_conv: upsampled external input");
  } else {
    if (int e = check_buf(p, d.src_buf, d.src_ch_off, d.cin, "rsb_plan_add_conv(src)")) return e;
    scale = p->bufs[d.src_buf].scale * (d.src_upsample2 ? 2 : 1);
  }
  if (d.dst_buf == RSB_EXTERNAL_OUTPUT) {
    if (d.ps < 1 || d.cout % (d.ps * d.ps) != 0 || d.cout / (d.ps * d.ps) != p->out_ch || scale * d.ps != p->upscale * p->base_div)
      return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: external output shape mismatch (cout %d, ps %d, scale %d)", d.cout, d.ps, scale);
    if (d.add_base && p->in_ch != p->out_ch) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: add_base needs in_ch == out_ch");
    if (d.add_base && scale != p->base_div) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_conv: add_base on an upsampled grid");
  } else {
    const int dps = d.dst_ps > 1 ? d.dst_ps : 1;
    int main_ch = d.cout;
    if (dps > 1 && d.dst_phase >= 0) {
      if (d.dst_phase >= dps * dps || d.cout % 8 != 0 || d.dst2_buf >= 0)
        return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: bad single-phase sub-pixel destination");
    } else if (dps > 1) {
      if (d.cout % (dps * dps) != 0 || (d.cout / (dps * dps)) % 8 != 0 || d.dst2_buf >= 0)
        return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: sub-pixel destination needs cout = %d * 8k channels and no dst2", dps * dps);
      main_ch = d.cout / (dps * dps);
    } else if (d.dst2_buf >= 0) {
      if (d.split_ch <= 0 || d.split_ch >= d.cout || d.split_ch % 8 != 0)
        return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: split_ch %d must be a multiple of 8 inside (0, cout)", d.split_ch);
      if (int e = check_buf(p, d.dst2_buf, d.dst2_ch_off, d.cout - d.split_ch, "rsb_plan_add_conv(dst2)")) return e;
      if (p->bufs[d.dst2_buf].scale != scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: dst2 grid mismatch");
      main_ch = d.split_ch;
    }
    if (int e = check_buf(p, d.dst_buf, d.dst_ch_off, main_ch, "rsb_plan_add_conv(dst)")) return e;
    if (p->bufs[d.dst_buf].scale != scale * dps) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: dst buffer grid scale %d != %d", p->bufs[d.dst_buf].scale, scale * dps);
    if (d.dst_buf == d.src_buf && d.kh * d.kw > 1) {
      const int a0 = d.src_ch_off, a1 = d.src_ch_off + ceil_div(d.cin, 16) * 16, b0 = d.dst_ch_off, b1 = d.dst_ch_off + d.cout;
      if (a0 < b1 && b0 < a1) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: spatial conv cannot run in place");
    }
  }
  if (d.combine != RSB_COMB_NONE) {
    if (int e = check_buf(p, d.res1_buf, d.res1_ch_off, d.cout, "rsb_plan_add_conv(res1)")) return e;
    if (p->bufs[d.res1_buf].scale != scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: res1 grid mismatch");
    if (d.combine == RSB_COMB_AXPY && d.res2_buf >= 0) {
      if (int e = check_buf(p, d.res2_buf, d.res2_ch_off, d.cout, "rsb_plan_add_conv(res2)")) return e;
      if (p->bufs[d.res2_buf].scale != scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: res2 grid mismatch");
    }
  }
  if (d.ln_fold) {
    if (d.kh != 1 || d.kw != 1 || d.src_buf < 0 || d.act == RSB_ACT_PRELU || d.dst_buf < 0)
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_conv: a LayerNorm fold needs a 1x1 conv from a buffer to a buffer, without PReLU");
    if (int e = check_buf(p, d.ln_stats_buf, 0, 8, "rsb_plan_add_conv(ln_stats)")) return e;
    if (p->bufs[d.ln_stats_buf].scale != scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: ln_stats grid mismatch");
    if (d.ln_fold == 2 && (p->dtype != RSB_BF16 || !(d.ln_eps > 0.0f)))
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_conv: raw LayerNorm sums (ln_fold = 2) need a bf16 plan and ln_eps > 0");
  }
  if (d.ln_out) {
    if (p->dtype != RSB_BF16 || d.kh != 1 || d.kw != 1 || d.src_buf < 0 || d.dst_buf < 0 || d.dst_ps > 1 || d.dst2_buf >= 0 || d.res2_buf >= 0 ||
        d.border_bias != nullptr || d.combine == RSB_COMB_SPAB_GATE || d.cout > 256)
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_conv: ln_out needs a bf16 plan and a 1x1 conv between buffers with a plain planar destination (cout <= 256)");
    if (int e = check_buf(p, d.ln_out_buf, 0, 8, "rsb_plan_add_conv(ln_out)")) return e;
    if (p->bufs[d.ln_out_buf].scale != scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_conv: ln_out grid mismatch");
  }
  c.scale = scale;
  const size_t wn = (size_t)d.cout * d.cin * d.kh * d.kw;
  c.w.assign(d.weight, d.weight + wn);
  if (d.bias) c.b.assign(d.bias, d.bias + d.cout);
  if (d.act == RSB_ACT_PRELU) c.slopes.assign(d.act_slopes, d.act_slopes + d.cout);
  if (d.border_bias) c.border.assign(d.border_bias, d.border_bias + (size_t)16 * d.cout);
  c.d.weight = nullptr, c.d.bias = nullptr, c.d.act_slopes = nullptr, c.d.border_bias = nullptr;
  c.cin_pad16 = ceil_div(d.cin, 16) * 16;
  c.npad = ceil_div(d.cout, 16) * 16;
  c.cpad32 = ceil_div(d.cout, 32) * 32;
  c.cin_planes = ceil_div(d.cin, 8);
  c.tc_src_buf = d.src_buf, c.tc_src_ch_off = d.src_ch_off, c.tc_cin = c.cin_pad16, c.tc_kh = d.kh, c.tc_kw = d.kw;
  if (p->dtype == RSB_BF16 && d.src_buf == RSB_EXTERNAL_INPUT && d.cin * d.kh * d.kw <= 256) {
    Buffer hb;
    c.pack_planar = d.kh * d.kw > 1 && c.cin_pad16 <= 64;
    c.pack_k = c.pack_planar ? c.cin_pad16 : ceil_div(d.cin * d.kh * d.kw, 16) * 16;
    hb.channels = c.pack_k, hb.planes = c.pack_k / 8, hb.scale = p->base_div;
    p->bufs.push_back(hb);
    c.pack_buf = (int)p->bufs.size() - 1;
    c.tc_src_buf = c.pack_buf, c.tc_src_ch_off = 0, c.tc_cin = c.pack_k;
    c.tc_kh = c.pack_planar ? d.kh : 1, c.tc_kw = c.pack_planar ? d.kw : 1;
  }
  p->convs.push_back(std::move(c));
  p->ops.push_back({0, (int)p->convs.size() - 1});
  return 0;
}

int rsb_plan_add_groupnorm(rsb_plan* p, const rsb_groupnorm_desc* desc) {
  if (!p || !desc) return fail(RSB_ERR_INVALID, "rsb_plan_add_groupnorm: NULL argument");
  if (p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_add_groupnorm: plan already finalized");
  const rsb_groupnorm_desc& d = *desc;
  if (d.groups < 1 || d.channels < 1 || d.channels % d.groups != 0 || (d.channels / d.groups) % 8 != 0)
    return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_groupnorm: channels per group must be a multiple of 8 (%d / %d)", d.channels, d.groups);
  if (!d.gamma || !d.beta) return fail(RSB_ERR_INVALID, "rsb_plan_add_groupnorm: gamma/beta NULL");
  if (int e = check_buf(p, d.src_buf, d.src_ch_off, d.channels, "rsb_plan_add_groupnorm(src)")) return e;
  if (int e = check_buf(p, d.dst_buf, d.dst_ch_off, d.channels, "rsb_plan_add_groupnorm(dst)")) return e;
  if (d.skip_buf >= 0)
    if (int e = check_buf(p, d.skip_buf, d.skip_ch_off, d.channels, "rsb_plan_add_groupnorm(skip)")) return e;
  GnOp g;
  g.d = d;
  g.gamma.assign(d.gamma, d.gamma + d.channels);
  g.beta.assign(d.beta, d.beta + d.channels);
  g.d.gamma = nullptr, g.d.beta = nullptr;
  g.scale = p->bufs[d.src_buf].scale;
  p->gns.push_back(std::move(g));
  p->ops.push_back({1, (int)p->gns.size() - 1});
  return 0;
}

int rsb_plan_add_op(rsb_plan* p, const rsb_op_desc* desc) {
  if (!p || !desc) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: NULL argument");
  if (p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_add_op: plan already finalized");
  const rsb_op_desc& d = *desc;
  if (d.kind < RSB_OP_LAYERNORM || d.kind > RSB_OP_CHAN_AFFINE) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: unknown kind %d", d.kind);
  if (d.channels < 1) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: bad channel count");
  if (d.kind == RSB_OP_DYSAMPLE) {
    const int g = d.i[0], s = d.i[1], oc = d.i[2];
    if (d.dst_buf != RSB_EXTERNAL_OUTPUT) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: DySample writes the external output (end_conv fused)");
    const bool proj = d.i[3] != 0;  // src = per-group projected maps (4 channels per group), end_conv already applied
    if (proj && (d.channels != 4 * g || g > 64)) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: projected DySample needs 4 channels per group");
    if (g < 1 || s < 1 || d.channels % g != 0 || d.channels > 256 || 2 * g * s * s > 256 || oc < 1 || oc > 4 || oc != p->out_ch)
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: DySample needs channels %% groups == 0, channels <= 256, 2*groups*s^2 <= 256, out channels <= 4");
    if (int e = check_buf(p, d.src_buf, d.src_ch_off, d.channels, "rsb_plan_add_op(DySample src)")) return e;
    if (d.i[4] != 0 && d.i[4] < 2 * g * s * s) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: DySample gate offset overlaps the offsets");
    if (int e = check_buf(p, d.src2_buf, d.src2_ch_off, (d.i[4] > 0 ? d.i[4] : 0) + 2 * g * s * s, "rsb_plan_add_op(DySample offsets)")) return e;
    if (p->bufs[d.src_buf].scale != p->bufs[d.src2_buf].scale || p->bufs[d.src_buf].scale * s != p->upscale * p->base_div)
      return fail(RSB_ERR_INVALID, "rsb_plan_add_op: DySample grid mismatch (buffer scale %d x %d != upscale %d)", p->bufs[d.src_buf].scale, s, p->upscale);
    if (!d.w[0] || d.wn[0] != 2 * g * s * s || !d.w[1] || (!proj && d.wn[1] != (int64_t)oc * d.channels) || !d.w[2] || d.wn[2] != oc)
      return fail(RSB_ERR_INVALID, "rsb_plan_add_op: DySample needs init_pos [2*groups*s^2], end_conv weight [out][C] and bias [out]");
    AuxOp a;
    a.d = d;
    for (int k = 0; k < 3; ++k) a.w[k].assign(d.w[k], d.w[k] + d.wn[k]);
    for (int k = 0; k < 8; ++k) a.d.w[k] = nullptr;
    a.scale = p->bufs[d.src_buf].scale;
    p->auxs.push_back(std::move(a));
    p->ops.push_back({2, (int)p->auxs.size() - 1});
    return 0;
  }
  if (d.kind == RSB_OP_UNSHUFFLE_POOL || d.kind == RSB_OP_SE_SHUFFLE) {
    const bool un = d.kind == RSB_OP_UNSHUFFLE_POOL;
    if (d.src_buf < 0 || d.src_buf >= (int)p->bufs.size() || d.dst_buf < 0 || d.dst_buf >= (int)p->bufs.size())
      return fail(RSB_ERR_INVALID, "rsb_plan_add_op: unknown buffer");
    if (d.channels % (un ? 8 : 32) != 0) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: %s needs channels %% %d == 0", un ? "unshuffle" : "shuffle", un ? 8 : 32);
    if (int e = check_buf(p, d.src_buf, d.src_ch_off, d.channels, "rsb_plan_add_op(src)")) return e;
    if (int e = check_buf(p, d.dst_buf, d.dst_ch_off, un ? 5 * d.channels : d.channels / 4, "rsb_plan_add_op(dst)")) return e;
    const int ss = p->bufs[d.src_buf].scale, ds = p->bufs[d.dst_buf].scale;
    if (un ? (ss != 2 * ds) : (ds != 2 * ss)) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: %s changes the grid by a factor of two (src scale %d, dst scale %d)", un ? "unshuffle" : "shuffle", ss, ds);
    if (!un && d.i[0] > 0) {
      const int hd = d.i[0];
      if (hd > 1024 || !d.w[0] || d.wn[0] != (int64_t)hd * d.channels || !d.w[1] || d.wn[1] != hd || !d.w[2] || d.wn[2] != (int64_t)hd * d.channels || !d.w[3] ||
          d.wn[3] != d.channels)
        return fail(RSB_ERR_INVALID, "rsb_plan_add_op: SE gate needs W1 [hidden][C], b1 [hidden], W2 [C][hidden], b2 [C]");
    }
    AuxOp a;
    a.d = d;
    for (int k = 0; k < 8; ++k) {
      if (d.w[k] && d.wn[k] > 0) a.w[k].assign(d.w[k], d.w[k] + d.wn[k]);
      a.d.w[k] = nullptr;
    }
    a.scale = ss;
    p->auxs.push_back(std::move(a));
    p->ops.push_back({2, (int)p->auxs.size() - 1});
    return 0;
  }
  if (d.kind == RSB_OP_DWCONV3 && d.i[1] != 0 && d.i[1] != 3) {
    const int K = d.i[1];
    if ((K != 5 && K != 7 && K != 9 && K != 11) || d.i[0] != RSB_ACT_NONE || d.src2_buf >= 0 || !d.w[0] || d.wn[0] != (int64_t)d.channels * K * K || !d.w[1] || d.wn[1] != d.channels)
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: depthwise K x K needs K in {5, 7, 9, 11}, no activation / gate, weight [C][K*K], bias [C]");
  }
  if (d.kind == RSB_OP_CHAN_GATE && (d.channels % 8 != 0 || d.channels > 1024 || d.src2_buf < 0 || !d.w[0] || d.wn[0] != (int64_t)d.channels * d.channels || !d.w[1] ||
                                     d.wn[1] != d.channels || !d.w[2] || d.wn[2] != d.channels))
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: channel gate needs C %% 8 == 0, C <= 1024, src2, W [C][C], bias [C] and gamma [C]");
  if (d.kind == RSB_OP_CHAN_AFFINE && (d.channels % 8 != 0 || !d.w[0] || d.wn[0] != d.channels))
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: channel affine needs C %% 8 == 0 and a scale [C]");
  if (d.kind == RSB_OP_RMSNORM && (!d.w[0] || d.wn[0] != d.channels || !d.w[1] || d.wn[1] != d.channels))
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: RMSNorm needs scale [C] and offset [C]");
  const bool qkv = d.kind == RSB_OP_WINATTN || d.kind == RSB_OP_CHANATTN;
  if (d.src_buf < 0 || d.src_buf >= (int)p->bufs.size() || d.dst_buf < 0 || d.dst_buf >= (int)p->bufs.size())
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: unknown buffer");
  const int stride = d.kind == RSB_OP_WINATTN ? d.i[4] : (d.kind == RSB_OP_CHANATTN ? d.i[1] : 0);
  const int head_pad = d.kind == RSB_OP_WINATTN ? d.i[5] : 0;  // heads on 32-channel boundaries (tcgen05 window attention)
  const int qkv_span = head_pad ? d.i[0] * head_pad : d.channels;
  const int src_need = d.src_ch_off + (qkv ? 2 * (stride > 0 ? stride : d.channels) : 0) + qkv_span;
  const int dst_need = (d.kind == RSB_OP_LAYERNORM && d.i[0] >= 1) ? 8 : qkv_span;  // statistics mode writes one pixel chunk
  if (src_need > p->bufs[d.src_buf].planes * 8 || d.dst_ch_off + dst_need > p->bufs[d.dst_buf].planes * 8)
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: channel range exceeds buffer");
  if (!qkv && (d.src_ch_off % 8 != 0 || d.dst_ch_off % 8 != 0))
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: channel offsets must be multiples of 8");
  if (p->bufs[d.src_buf].scale != p->bufs[d.dst_buf].scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: grid mismatch");
  if (d.src2_buf >= 0) {
    if (int e = check_buf(p, d.src2_buf, d.src2_ch_off, d.channels, "rsb_plan_add_op(src2)")) return e;
  } else if (d.kind == RSB_OP_AIM) {
    return fail(RSB_ERR_INVALID, "rsb_plan_add_op: AIM needs src2 (the conv branch)");
  }
  if (qkv) {
    const int heads = d.i[0];
    if (heads < 1 || d.channels % heads != 0 || d.channels / heads > 32)
      return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: attention needs head_dim <= 32 (dim %d, heads %d)", d.channels, heads);
    if (d.kind == RSB_OP_WINATTN) {
      if (heads % 2 != 0 || d.i[1] < 1 || d.i[2] < 1 || d.i[1] * d.i[2] > 256)
        return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: window attention needs even heads and <= 256 tokens per window");
      const int64_t tab = (int64_t)(2 * d.i[1] - 1) * (2 * d.i[2] - 1) * (heads / 2);
      if (!d.w[0] || !d.w[1] || d.wn[0] != tab || d.wn[1] != tab)
        return fail(RSB_ERR_INVALID, "rsb_plan_add_op: position-bias tables must hold %lld entries", (long long)tab);
      if (head_pad != 0) {
        if (head_pad != 32 || p->dtype != RSB_BF16 || !rsb::winattn_tc_supported(heads, d.channels / heads, d.i[1], d.i[2]))
          return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: head-padded window attention (i[5] = 32) needs a bf16 plan, head_dim < 32 and "
                                           "8-aligned windows of 64 / 128 / 256 tokens");
        if (d.src_ch_off % 8 != 0 || d.dst_ch_off % 8 != 0 || stride % 8 != 0 || stride < heads * head_pad)
          return fail(RSB_ERR_INVALID, "rsb_plan_add_op: head-padded window attention needs 8-aligned channel offsets and a q/k/v stride >= heads * 32");
      }
    }
  }
  if (d.kind == RSB_OP_AIM && (d.i[1] < 1 || d.i[1] > 64 || d.i[2] < 1 || d.i[2] > 16 || d.channels > 512))
    return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: AIM hidden widths out of range");
  if (d.kind == RSB_OP_AIM && d.i[3] != 0) {
    if (p->dtype != RSB_BF16) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_add_op: AIM with a precomputed hidden map is a bf16-plan feature");
    if (int e = check_buf(p, d.i[3] - 1, 0, 16, "rsb_plan_add_op(AIM hidden)")) return e;
    if (p->bufs[d.i[3] - 1].scale != p->bufs[d.src_buf].scale) return fail(RSB_ERR_INVALID, "rsb_plan_add_op: AIM hidden map grid mismatch");
  }
  AuxOp a;
  a.d = d;
  for (int k = 0; k < 8; ++k) {
    if (d.w[k] && d.wn[k] > 0) a.w[k].assign(d.w[k], d.w[k] + d.wn[k]);
    a.d.w[k] = nullptr;
  }
  a.scale = p->bufs[d.src_buf].scale;
  p->auxs.push_back(std::move(a));
  p->ops.push_back({2, (int)p->auxs.size() - 1});
  return 0;
}

int rsb_plan_finalize(rsb_plan* p, int device) {
  if (!p) return fail(RSB_ERR_INVALID, "rsb_plan_finalize: NULL plan");
  if (p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_finalize: already finalized");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return fail(RSB_ERR_NO_DEVICE, "rsb_plan_finalize: CUDA device %d not available (%d visible)", device, count);
  }
  cudaDeviceProp prop;
  RSB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RSB_ERR_NO_DEVICE, "rsb_plan_finalize: device %d is sm_%d%d; this library only contains sm_100a code", device, prop.major, prop.minor);
  DeviceGuard guard(device);
  p->device = device;
  p->num_sms = prop.multiProcessorCount;
  RSB_CUDA(rsb::conv_tc_configure(kMaxSmem));
  RSB_CUDA(rsb::conv_rs_configure(kMaxSmem));
  RSB_CUDA(rsb::conv_pair_configure(kMaxSmem));
  RSB_CUDA(rsb::conv_lk_configure(kMaxSmem));
  static const bool no_rs = rsb::rsb_env("RSB_NO_RS") != nullptr;

  for (ConvOp& c : p->convs) {
    const rsb_conv_desc& d = c.d;
    const int taps = d.kh * d.kw;
    // ---- tensor-core eligibility (bf16 plan, planar source, operands fit shared memory)
    c.tc_ok = false;
    if (p->dtype == RSB_BF16 && c.tc_src_buf >= 0 && !d.src_upsample2 && c.npad <= 256) {
      const Buffer& sb = p->bufs[c.tc_src_buf];
      const bool planes_ok = c.tc_src_ch_off / 8 + c.tc_cin / 8 <= sb.planes;
      const int WT = rsb::kTileW + c.tc_kw - 1, HT = rsb::kTileH + c.tc_kh - 1;
      // prefer the whole Cin per stage (static-geometry kernels); otherwise stage K chunks of 64/48/32/16 channels.
      // Up to 8 stages: a 48-channel 1x1 conv stages only 12 KB per tile, and four of those in flight per SM cap the kernel at
      // ~2 TB/s (bytes in flight = bandwidth x latency); big stages still get as many as fit.
      static const int kMaxTcStages = rsb::rsb_env("RSB_TC_STAGES4") ? 4 : 8;
      int stages = 0, kchunk = 0;
      const int cands[5] = {c.tc_cin, 64, 48, 32, 16};
      for (int ci = 0; ci < 5 && stages == 0; ++ci) {
        const int kc = cands[ci];
        if (kc > c.tc_cin || c.tc_cin % kc != 0) continue;
        const int min_stages = ci == 0 ? 2 : 3;
        for (int s = kMaxTcStages; s >= min_stages; --s)
          if (rsb::conv_tc_smem_bytes(c.tc_cin, kc, c.npad, c.tc_kh, c.tc_kw, s) <= kMaxSmem) {
            stages = s, kchunk = kc;
            break;
          }
      }
      if (planes_ok && stages >= 2 && 8 * WT <= 256 && HT <= 256 && HT * WT < 16384) {
        c.tc_ok = true;
        c.stages = stages;
        c.kchunk = kchunk;
      }
      if (d.ln_out && !c.tc_ok) return fail(RSB_ERR_UNSUPPORTED, "rsb_plan_finalize: a conv with ln_out does not fit the tensor-core tile kernel");
    }
    const int cmax = std::max(c.npad, c.cpad32);
    std::vector<float> bias(cmax, 0.0f), slopes(cmax, 0.0f);
    for (int o = 0; o < d.cout; ++o) {
      if (!c.b.empty()) bias[o] = c.b[o];
      if (!c.slopes.empty()) slopes[o] = c.slopes[o];
    }
    RSB_CUDA(cudaMalloc(&c.d_bias, cmax * sizeof(float)));
    RSB_CUDA(cudaMemcpy(c.d_bias, bias.data(), cmax * sizeof(float), cudaMemcpyHostToDevice));
    if (!c.border.empty()) {
      std::vector<float> bb((size_t)16 * cmax, 0.0f);
      for (int m = 0; m < 16; ++m)
        for (int o = 0; o < d.cout; ++o) bb[(size_t)m * cmax + o] = c.border[(size_t)m * d.cout + o];
      RSB_CUDA(cudaMalloc(&c.d_border, bb.size() * sizeof(float)));
      RSB_CUDA(cudaMemcpy(c.d_border, bb.data(), bb.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (d.ln_fold) {
      // row sums of the weights exactly as the kernels multiply them (bf16 plans round the weights)
      std::vector<float> rs(cmax, 0.0f);
      for (int o = 0; o < d.cout; ++o) {
        double acc = 0.0;
        for (int k = 0; k < d.cin; ++k) {
          float w = c.w[(size_t)o * d.cin + k];
          if (p->dtype == RSB_BF16) {
            const uint32_t u = (uint32_t)f32_to_bf16(w) << 16;
            memcpy(&w, &u, 4);
          }
          acc += (double)w;
        }
        rs[o] = (float)acc;
      }
      RSB_CUDA(cudaMalloc(&c.d_lnsum, cmax * sizeof(float)));
      RSB_CUDA(cudaMemcpy(c.d_lnsum, rs.data(), cmax * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (d.act == RSB_ACT_PRELU) {
      RSB_CUDA(cudaMalloc(&c.d_slopes, cmax * sizeof(float)));
      RSB_CUDA(cudaMemcpy(c.d_slopes, slopes.data(), cmax * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (c.tc_ok && c.pack_buf >= 0 && !c.pack_planar) {
      // packed conv == 1x1 conv over k = (ci*kh + ky)*kw + kx, which is the OIHW flattening of the weight
      const int k8 = c.pack_k / 8, kreal = d.cin * taps;
      std::vector<uint16_t> wp((size_t)k8 * c.npad * 8, 0);
      for (int o = 0; o < d.cout; ++o)
        for (int k = 0; k < kreal; ++k) wp[((size_t)(k / 8) * c.npad + o) * 8 + (k & 7)] = f32_to_bf16(c.w[(size_t)o * kreal + k]);
      c.wbytes_tc = (uint32_t)(wp.size() * 2);
      RSB_CUDA(cudaMalloc(&c.d_wtc, c.wbytes_tc));
      RSB_CUDA(cudaMemcpy(c.d_wtc, wp.data(), c.wbytes_tc, cudaMemcpyHostToDevice));
    } else if (c.tc_ok) {
      // [tap][cin/8][npad][8] bf16: per (tap, 8-channel slab) an N x 8 K-major panel of 128-byte core matrices
      const int cin8 = c.cin_pad16 / 8;
      std::vector<uint16_t> wp((size_t)taps * cin8 * c.npad * 8, 0);
      for (int o = 0; o < d.cout; ++o)
        for (int ci = 0; ci < d.cin; ++ci)
          for (int t = 0; t < taps; ++t)
            wp[(((size_t)t * cin8 + ci / 8) * c.npad + o) * 8 + (ci & 7)] = f32_to_bf16(c.w[((size_t)o * d.cin + ci) * taps + t]);
      c.wbytes_tc = (uint32_t)(wp.size() * 2);
      RSB_CUDA(cudaMalloc(&c.d_wtc, c.wbytes_tc));
      RSB_CUDA(cudaMemcpy(c.d_wtc, wp.data(), c.wbytes_tc, cudaMemcpyHostToDevice));
      if (!no_rs && d.kh == 3 && d.kw == 3 && d.pad_t == 1 && d.pad_l == 1 && 3 * c.npad <= 256 && 512 / c.npad >= 5) {
        // row-streaming kernel: [kw][cin/8][kh * npad + o][8] — the three kernel rows side by side on the N axis
        // input-ring depth: 6 / 8 / 12 stages measured 2.09 / 2.07 / 2.05 ms on SPAN 1080p (bring-up builds: RSB_RS_STAGES, clamped
        // to the range the barrier layout was validated with)
        static const int kMaxRsStages = rsb::rsb_env("RSB_RS_STAGES") ? std::min(12, std::max(3, atoi(rsb::rsb_env("RSB_RS_STAGES")))) : 8;
        for (int s = kMaxRsStages; s >= 3 && c.rs_stages == 0; --s)
          if (rsb::conv_rs_smem_bytes(c.tc_cin, c.npad, s) <= kMaxSmem) c.rs_stages = s;
        if (c.rs_stages > 0) {
          const int n3 = 3 * c.npad;
          std::vector<uint16_t> wr((size_t)3 * cin8 * n3 * 8, 0);
          for (int o = 0; o < d.cout; ++o)
            for (int ci = 0; ci < d.cin; ++ci)
              for (int ky = 0; ky < 3; ++ky)
                for (int kx = 0; kx < 3; ++kx)
                  wr[(((size_t)kx * cin8 + ci / 8) * n3 + ky * c.npad + o) * 8 + (ci & 7)] =
                      f32_to_bf16(c.w[((size_t)o * d.cin + ci) * 9 + ky * 3 + kx]);
          c.wbytes_rs = (uint32_t)(wr.size() * 2);
          RSB_CUDA(cudaMalloc(&c.d_wrs, c.wbytes_rs));
          RSB_CUDA(cudaMemcpy(c.d_wrs, wr.data(), c.wbytes_rs, cudaMemcpyHostToDevice));
          c.rs_elig = true;
        }
      }
      if (!no_rs && d.kh == d.kw && d.kh % 2 == 1 && d.kh >= 5 && d.kh <= 17 && d.pad_t == d.kh / 2 && d.pad_l == d.kw / 2 && c.npad == 16 &&
          c.tc_cin <= 64) {
        // large-kernel row-streaming kernel: [kw][cin/8][K * 16][8], N block j holds kernel row K-1-j
        const int K = d.kh, nk = K * 16;
        for (int sg = 8; sg >= 4 && c.lk_stages == 0; --sg)
          if (rsb::conv_lk_smem_bytes(c.tc_cin, K, sg) <= kMaxSmem) c.lk_stages = sg;
        if (c.lk_stages > 0) {
          std::vector<uint16_t> wl((size_t)K * cin8 * nk * 8, 0);
          for (int o = 0; o < d.cout; ++o)
            for (int ci = 0; ci < d.cin; ++ci)
              for (int j = 0; j < K; ++j)
                for (int kx = 0; kx < K; ++kx)
                  wl[(((size_t)kx * cin8 + ci / 8) * nk + j * 16 + o) * 8 + (ci & 7)] =
                      f32_to_bf16(c.w[((size_t)o * d.cin + ci) * K * K + (K - 1 - j) * K + kx]);
          c.wbytes_lk = (uint32_t)(wl.size() * 2);
          RSB_CUDA(cudaMalloc(&c.d_wlk, c.wbytes_lk));
          RSB_CUDA(cudaMemcpy(c.d_wlk, wl.data(), c.wbytes_lk, cudaMemcpyHostToDevice));
          c.lk_elig = true;
        }
      }
    }
    {
      // [cin_planes][kh][kw][8][cpad32] fp32 for the CUDA-core kernel
      std::vector<float> wp((size_t)c.cin_planes * taps * 8 * c.cpad32, 0.0f);
      for (int o = 0; o < d.cout; ++o)
        for (int ci = 0; ci < d.cin; ++ci)
          for (int t = 0; t < taps; ++t)
            wp[((((size_t)(ci / 8)) * taps + t) * 8 + (ci & 7)) * c.cpad32 + o] = c.w[((size_t)o * d.cin + ci) * taps + t];
      RSB_CUDA(cudaMalloc(&c.d_wdirect, wp.size() * sizeof(float)));
      RSB_CUDA(cudaMemcpy(c.d_wdirect, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    std::vector<float>().swap(c.w);
  }
  find_pairs(p);
  for (GnOp& g : p->gns) {
    RSB_CUDA(cudaMalloc(&g.d_gamma, g.gamma.size() * sizeof(float)));
    RSB_CUDA(cudaMalloc(&g.d_beta, g.beta.size() * sizeof(float)));
    RSB_CUDA(cudaMemcpy(g.d_gamma, g.gamma.data(), g.gamma.size() * sizeof(float), cudaMemcpyHostToDevice));
    RSB_CUDA(cudaMemcpy(g.d_beta, g.beta.data(), g.beta.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  RSB_CUDA(rsb::winattn_configure());
  for (AuxOp& a : p->auxs)
    for (int k = 0; k < 8; ++k)
      if (!a.w[k].empty()) {
        RSB_CUDA(cudaMalloc(&a.dw[k], a.w[k].size() * sizeof(float)));
        RSB_CUDA(cudaMemcpy(a.dw[k], a.w[k].data(), a.w[k].size() * sizeof(float), cudaMemcpyHostToDevice));
      }
  p->finalized = true;
  // A bf16 plan is meant to run on tensor cores; a conv that does not fit them runs on the CUDA-core kernel (~30 TFLOP/s).
  // Correct, but a performance cliff the caller should know about: count them (rsb_plan_num_direct_convs) and say so once.
  p->num_direct = 0;
  if (p->dtype == RSB_BF16)
    for (const ConvOp& c : p->convs) p->num_direct += c.tc_ok ? 0 : 1;
  if (p->num_direct > 0) {
    static std::once_flag warned;
    const int nd = p->num_direct, nc = (int)p->convs.size();
    std::call_once(warned, [nd, nc] {
      fprintf(stderr, "resselt_b200: %d of %d convolutions of a bf16 plan do not fit the tensor-core kernels and run on the CUDA-core "
                      "kernel (see rsb_plan_num_direct_convs)\n", nd, nc);
    });
  }
  return 0;
}

int rsb_plan_num_ops(const rsb_plan* p) { return p ? (int)p->ops.size() : 0; }
int rsb_plan_num_direct_convs(const rsb_plan* p) { return p ? p->num_direct : 0; }
int rsb_plan_set_nvtx(rsb_plan* p, int enable) {
  if (!p) return fail(RSB_ERR_INVALID, "rsb_plan_set_nvtx: NULL plan");
  p->nvtx = enable != 0;
  return 0;
}

const char* rsb_kernel_name(int k) {
  static const char* const names[] = {"conv_direct", "conv_tc", "conv_rs", "conv_lk", "conv_pair", "groupnorm", "layernorm", "dwconv3",
                                      "winattn", "chanattn", "aim", "dysample", "rmsnorm", "unshuffle_pool", "se_shuffle", "chan_gate", "chan_affine"};
  return (k >= 0 && k < (int)(sizeof(names) / sizeof(names[0]))) ? names[k] : "unknown";
}

int rsb_plan_op_info(const rsb_plan* p, int op_index, rsb_op_info* out) {
  if (!p || !out) return fail(RSB_ERR_INVALID, "rsb_plan_op_info: NULL argument");
  if (op_index < 0 || op_index >= (int)p->ops.size()) return fail(RSB_ERR_INVALID, "rsb_plan_op_info: op %d outside [0, %zu)", op_index, p->ops.size());
  if (!p->bws) return fail(RSB_ERR_STATE, "rsb_plan_op_info: no forward has bound a shape yet");
  memset(out, 0, sizeof *out);
  const Op& op = p->ops[op_index];
  out->kind = op.kind;
  out->launches = 1;
  const double es = (double)p->elem();
  if (op.kind == 0) {
    const ConvOp& c = p->convs[op.index];
    auto px = [&](const ConvOp& k) { return (double)p->bn * (p->bh * k.scale) * (double)(p->bw * k.scale); };
    auto flops = [&](const ConvOp& k) { return 2.0 * k.d.cout * k.d.cin * k.d.kh * k.d.kw * px(k); };
    auto out_bytes = [&](const ConvOp& k) { return (double)k.d.cout * px(k) * es; };
    auto res_bytes = [&](const ConvOp& k) {
      return k.d.combine == RSB_COMB_NONE ? 0.0 : (double)k.d.cout * px(k) * es * (k.d.combine == RSB_COMB_AXPY && k.d.res2_buf >= 0 ? 2 : 1);
    };
    auto in_bytes = [&](const ConvOp& k) { return (double)(k.d.src_buf == RSB_EXTERNAL_INPUT ? k.d.cin : k.tc_cin) * px(k) * es; };
    // fused pairs are the opt-in mode 4: pair_ready says whether op_index heads one
    const bool fused_mode = p->info_mode == 4;
    if (fused_mode && op_index > 0 && p->ops[op_index - 1].kind == 0 && p->convs[p->ops[op_index - 1].index].pair_ready) {
      out->kernel = RSB_K_CONV_PAIR, out->launches = 0;
      return 0;
    }
    if (fused_mode && c.pair_ready && op_index + 1 < (int)p->ops.size()) {
      const ConvOp& b = p->convs[p->ops[op_index + 1].index];
      out->kernel = RSB_K_CONV_PAIR, out->fused_next = 1;
      out->launches = 1 + (c.pack_buf >= 0 ? 1 : 0);
      out->flops = flops(c) + flops(b);
      out->bytes = in_bytes(c) + res_bytes(c) + (c.pair_store ? out_bytes(c) : 0.0) + res_bytes(b) + out_bytes(b);
      return 0;
    }
    out->flops = flops(c);
    out->bytes = in_bytes(c) + res_bytes(c) + out_bytes(c);
    if (p->info_mode != 1 && c.grouped) {  // runs inside the launch of the group's head
      out->kernel = RSB_K_CONV_TC, out->launches = 0, out->flops = 0.0, out->bytes = 0.0;
      return 0;
    }
    if (p->info_mode != 1 && c.group_n > 1 && op_index + c.group_n <= (int)p->ops.size()) {
      out->kernel = RSB_K_CONV_TC, out->fused_next = c.group_n - 1;
      for (int j = 1; j < c.group_n; ++j) {
        const ConvOp& b = p->convs[p->ops[op_index + j].index];
        out->flops += flops(b), out->bytes += out_bytes(b);  // the source is read once
      }
      return 0;
    }
    if (!c.tc_ok)
      out->kernel = RSB_K_CONV_DIRECT;
    else {
      out->launches = 1 + (c.pack_buf >= 0 ? 1 : 0);
      out->kernel = (c.lk_ready && c.rs_pref) ? RSB_K_CONV_LK : ((c.rs_ready && c.rs_pref) ? RSB_K_CONV_RS : RSB_K_CONV_TC);
    }
  } else if (op.kind == 1) {
    out->kernel = RSB_K_GROUPNORM, out->launches = 2;
  } else {
    switch (p->auxs[op.index].d.kind) {
      case RSB_OP_LAYERNORM: out->kernel = RSB_K_LAYERNORM; break;
      case RSB_OP_DWCONV3: out->kernel = RSB_K_DWCONV3; break;
      case RSB_OP_WINATTN: out->kernel = RSB_K_WINATTN; break;
      case RSB_OP_CHANATTN: out->kernel = RSB_K_CHANATTN, out->launches = 3; break;
      case RSB_OP_AIM: out->kernel = RSB_K_AIM, out->launches = 3; break;
      case RSB_OP_RMSNORM: out->kernel = RSB_K_RMSNORM; break;
      case RSB_OP_UNSHUFFLE_POOL: out->kernel = RSB_K_UNSHUFFLE_POOL; break;
      case RSB_OP_SE_SHUFFLE: out->kernel = RSB_K_SE_SHUFFLE, out->launches = p->auxs[op.index].d.i[0] > 0 ? 3 : 1; break;
      case RSB_OP_CHAN_GATE: out->kernel = RSB_K_CHAN_GATE, out->launches = 3; break;
      case RSB_OP_CHAN_AFFINE: out->kernel = RSB_K_CHAN_AFFINE; break;
      default: out->kernel = RSB_K_DYSAMPLE; break;
    }
  }
  return 0;
}

int rsb_plan_launches_per_forward(const rsb_plan* p) {
  if (!p) return 0;
  int packed = 0;
  for (const ConvOp& c : p->convs) packed += (c.pack_buf >= 0 && c.tc_ok) ? 1 : 0;
  int aux = 0;
  for (const AuxOp& a : p->auxs) aux += (a.d.kind == RSB_OP_CHANATTN || a.d.kind == RSB_OP_AIM || a.d.kind == RSB_OP_CHAN_GATE || (a.d.kind == RSB_OP_SE_SHUFFLE && a.d.i[0] > 0)) ? 3 : 1;
  return (int)p->convs.size() + packed + 2 * (int)p->gns.size() + aux;  // mode 0; mode 4 saves one launch per fused pair
}

int rsb_plan_flops(const rsb_plan* p, int n, int h, int w, double* flops) {
  if (!p || !flops) return fail(RSB_ERR_INVALID, "rsb_plan_flops: NULL argument");
  double f = 0.0;
  h /= p->base_div, w /= p->base_div;
  for (const ConvOp& c : p->convs)
    f += 2.0 * c.d.cout * c.d.cin * c.d.kh * c.d.kw * (double)n * (h * c.scale) * (double)(w * c.scale);
  *flops = f;
  return 0;
}

int rsb_plan_workspace_bytes(const rsb_plan* p, int n, int h, int w, size_t* bytes) {
  if (!p || !bytes) return fail(RSB_ERR_INVALID, "rsb_plan_workspace_bytes: NULL argument");
  if (n < 1 || h < 1 || w < 1) return fail(RSB_ERR_INVALID, "rsb_plan_workspace_bytes: bad shape");
  if (h % p->base_div || w % p->base_div) return fail(RSB_ERR_INVALID, "rsb_plan_workspace_bytes: %d x %d is not a multiple of the plan's base divisor %d", h, w, p->base_div);
  return layout(p, n, h / p->base_div, w / p->base_div, nullptr, nullptr, bytes);
}

int rsb_plan_forward(rsb_plan* p, const void* x, int x_dtype, int n, int h, int w, void* y, int y_dtype, void* workspace,
                     size_t workspace_bytes, void* stream_, int force_direct) {
  return rsb_plan_forward_ops(p, x, x_dtype, n, h, w, y, y_dtype, workspace, workspace_bytes, stream_, force_direct, 0,
                              p ? (int)p->ops.size() : 0);
}

int rsb_plan_forward_ops(rsb_plan* p, const void* x, int x_dtype, int n, int h, int w, void* y, int y_dtype, void* workspace,
                         size_t workspace_bytes, void* stream_, int force_direct, int op_begin, int op_end) {
  if (!p || !x || !y || !workspace) return fail(RSB_ERR_INVALID, "rsb_plan_forward: NULL argument");
  if (op_begin < 0 || op_end > (int)p->ops.size() || op_begin > op_end)
    return fail(RSB_ERR_INVALID, "rsb_plan_forward_ops: op range [%d, %d) outside [0, %zu)", op_begin, op_end, p->ops.size());
  if (!p->finalized) return fail(RSB_ERR_STATE, "rsb_plan_forward: plan not finalized");
  if (n < 1 || h < 1 || w < 1) return fail(RSB_ERR_INVALID, "rsb_plan_forward: bad shape %dx%dx%d", n, h, w);
  if (x_dtype < RSB_F32 || x_dtype > RSB_F16 || y_dtype < RSB_F32 || y_dtype > RSB_F16)
    return fail(RSB_ERR_INVALID, "rsb_plan_forward: bad tensor dtype");
  if (h % p->base_div || w % p->base_div) return fail(RSB_ERR_INVALID, "rsb_plan_forward: %d x %d is not a multiple of the plan's base divisor %d", h, w, p->base_div);
  h /= p->base_div, w /= p->base_div;  // everything below works on the plan's base grid
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  std::lock_guard<std::mutex> lock(p->mu);
  DeviceGuard guard(p->device);
  p->info_mode = force_direct;
  int rc = 0;
  if (p->bn != n || p->bh != h || p->bw != w || p->bws != workspace) rc = bind(p, n, h, w, workspace, workspace_bytes, stream);
  if (rc == 0) {
    for (int oi = op_begin; oi < op_end; ++oi) {
      const Op& op = p->ops[oi];
      cudaError_t e;
      rsb_op_info info;
      if (p->nvtx) rsb_plan_op_info(p, oi, &info);
      NvtxRange range(p->nvtx, p->nvtx ? rsb_kernel_name(info.kernel) : "", oi);
      if (op.kind == 0) {
        ConvOp& c = p->convs[op.index];
        if (c.pair_ready && force_direct == 4 && oi + 1 < op_end) {
          // opt-in (mode 4): this conv and the next one as one fused launch; the intermediate map is not re-read from HBM.
          // Not the default: measured on B200 the fused pair is bound by shared-memory bandwidth (operand reads of the N = 144 MMAs
          // + ring / stage traffic) at about the time the two HBM-bound single launches take (profiles/r2_conv_pair_analysis.md)
          if (c.pack_buf >= 0) {
            rsb::PackParams k = c.pk;
            k.src = x, k.src_dtype = x_dtype;
            e = rsb::launch_pack_input(k, stream);
            if (e != cudaSuccess) {
              rc = fail_cuda(e, "kernel launch");
              break;
            }
          }
          rsb::ConvPairParams q = c.prp;
          e = rsb::launch_conv_pair(c.map_rs, q.res_prefetch ? c.map_res : c.map_rs, q, p->num_sms, stream);
          ++oi;
        } else if (c.tc_ok && force_direct != 1) {
          if (c.pack_buf >= 0) {
            rsb::PackParams k = c.pk;
            k.src = x, k.src_dtype = x_dtype;
            e = rsb::launch_pack_input(k, stream);
            if (e != cudaSuccess) {
              rc = fail_cuda(e, "kernel launch");
              break;
            }
          }
          rsb::ConvTcParams t = c.tcp;
          if (t.epi.dst_external) t.epi.dst = y, t.epi.out_dtype = y_dtype;
          t.epi.base = x, t.epi.base_dtype = x_dtype;
          if (c.lk_ready && force_direct != 2 && (c.rs_pref || force_direct == 3)) {
            rsb::ConvLkParams q = c.lkp;
            q.epi = t.epi;
            e = rsb::launch_conv_lk(c.map_rs, q, p->num_sms, stream);
          } else if (c.rs_ready && force_direct != 2 && (c.rs_pref || force_direct == 3)) {
            rsb::ConvRsParams q = c.rsp;
            q.epi = t.epi;
            e = rsb::launch_conv_rs(c.map_rs, q, p->num_sms, stream);
          } else if (c.group_n > 1 && oi + c.group_n <= op_end) {
            e = rsb::launch_conv_tc(c.map, c.gtcp, p->num_sms, stream);  // this conv and the next group_n - 1 as one launch
            oi += c.group_n - 1;
          } else
            e = rsb::launch_conv_tc(c.map, t, p->num_sms, stream);
        } else {
          rsb::ConvDirectParams q = c.dp;
          if (q.src_external) q.src = x, q.src_dtype = x_dtype;
          if (q.epi.dst_external) q.epi.dst = y, q.epi.out_dtype = y_dtype;
          q.epi.base = x, q.epi.base_dtype = x_dtype;
          e = rsb::launch_conv_direct(q, p->dtype == RSB_BF16, stream);
          if (e == cudaSuccess && c.d.ln_out)  // the CUDA-core kernel has no ln_out epilogue: the raw sums from the map it wrote
            e = rsb::launch_ln_raw_sums(q.epi.dst, q.epi.dst_planes, q.epi.dst_plane0, ceil_div(c.d.cout, 8), c.tcp.n, q.epi.H, q.epi.W, c.tcp.epi.ln_out,
                                        c.tcp.epi.ln_out_planes, p->num_sms, stream);
        }
      } else if (op.kind == 1) {
        e = rsb::launch_groupnorm(p->gns[op.index].gp, p->dtype == RSB_BF16, stream);
      } else {
        AuxOp& a = p->auxs[op.index];
        const bool bf = p->dtype == RSB_BF16;
        switch (a.d.kind) {
          case RSB_OP_LAYERNORM: e = rsb::launch_layernorm(a.tok, bf, p->num_sms, stream); break;
          case RSB_OP_DWCONV3:
            e = (a.d.i[1] > 3) ? rsb::launch_dwconv_k(a.tok, a.d.i[1], bf, p->num_sms, stream) : rsb::launch_dwconv3(a.tok, bf, stream);
            break;
          case RSB_OP_RMSNORM: e = rsb::launch_rmsnorm(a.tok, bf, p->num_sms, stream); break;
          case RSB_OP_UNSHUFFLE_POOL: e = rsb::launch_unshuffle_pool(a.tok, bf, p->num_sms, stream); break;
          case RSB_OP_SE_SHUFFLE: e = rsb::launch_se_shuffle(a.se, bf, p->num_sms, stream); break;
          case RSB_OP_CHAN_GATE: e = rsb::launch_chan_gate(a.se, bf, p->num_sms, stream); break;
          case RSB_OP_CHAN_AFFINE: e = rsb::launch_chan_affine(a.tok, bf, p->num_sms, stream); break;
          case RSB_OP_WINATTN: e = rsb::launch_winattn(a.win, bf, p->num_sms, stream); break;
          case RSB_OP_CHANATTN: e = rsb::launch_chanattn(a.chan, bf, p->num_sms, stream); break;
          case RSB_OP_DYSAMPLE: {
            rsb::DySampleParams q = a.dys;
            q.dst = y, q.dst_dtype = y_dtype;
            e = rsb::launch_dysample(q, bf, stream);
            break;
          }
          default: e = rsb::launch_aim(a.aim, bf, p->num_sms, stream); break;
        }
      }
      if (e != cudaSuccess) {
        rc = fail_cuda(e, "kernel launch");
        break;
      }
    }
  }
  return rc;
}

int rsb_plan_read_buffer(rsb_plan* p, int buf_id, int ch_off, int channels, float* dst_nchw, void* stream_) {
  if (!p || !dst_nchw) return fail(RSB_ERR_INVALID, "rsb_plan_read_buffer: NULL argument");
  if (!p->bws) return fail(RSB_ERR_STATE, "rsb_plan_read_buffer: no forward has run yet");
  if (int e = check_buf(p, buf_id, ch_off, channels, "rsb_plan_read_buffer")) return e;
  const Buffer& b = p->bufs[buf_id];
  cudaError_t e = rsb::launch_planar_to_nchw(reinterpret_cast<uint8_t*>(p->bws) + b.offset, p->dtype == RSB_BF16, p->bn, b.planes,
                                             ch_off / 8, channels, p->bh * b.scale, p->bw * b.scale, dst_nchw,
                                             reinterpret_cast<cudaStream_t>(stream_));
  if (e != cudaSuccess) return fail_cuda(e, "planar_to_nchw");
  return 0;
}

}  // extern "C"
