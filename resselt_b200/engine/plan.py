"""Python face of a native plan: the layer program an architecture plugin emits.

``PlanBuilder`` records buffers and fused ops through the C ABI (include/resselt_b200.h);
``Plan`` owns the finalised native handle and replays it on tensors.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import weakref
from collections import OrderedDict
from typing import Optional, Sequence

import numpy as np
import torch

from . import native as N

_log = logging.getLogger('resselt_b200')

_TORCH_TO_RSB = {torch.float32: N.F32, torch.bfloat16: N.BF16, torch.float16: N.F16}


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().to('cpu', torch.float64).numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Ref:
    """A channel range [ch_off, ch_off + channels) of a plan buffer."""

    __slots__ = ('buf', 'ch_off', 'channels')

    def __init__(self, buf: int, ch_off: int, channels: int):
        self.buf, self.ch_off, self.channels = buf, ch_off, channels

    def slice(self, start: int, channels: int) -> 'Ref':
        assert start + channels <= self.channels
        return Ref(self.buf, self.ch_off + start, channels)


INPUT = Ref(N.EXTERNAL_INPUT, 0, 0)
OUTPUT = Ref(N.EXTERNAL_OUTPUT, 0, 0)


class PlanBuilder:
    def __init__(self, compute_dtype: torch.dtype, in_channels: int, out_channels: int, upscale: int, base_divisor: int = 1):
        self._lib = N.lib()
        handle = C.c_void_p()
        N.check(self._lib.rsb_plan_create(_TORCH_TO_RSB[compute_dtype], in_channels, out_channels, upscale, C.byref(handle)))
        self._h = handle
        # buffer grids are (H / base_divisor * scale): plans with a branch coarser than the input (RTMoSR) use 2, their full-resolution
        # buffers then have scale 2
        self.base_divisor = int(base_divisor)
        if self.base_divisor != 1:
            N.check(self._lib.rsb_plan_set_base_divisor(self._h, self.base_divisor))
        self.compute_dtype = compute_dtype
        self.in_channels, self.out_channels, self.upscale = in_channels, out_channels, upscale
        self.scales = {}

    def buffer(self, channels: int, scale: Optional[int] = None) -> Ref:
        """Activation buffer; ``scale`` defaults to the input's own grid (= base_divisor)."""
        scale = self.base_divisor if scale is None else scale
        bid = C.c_int()
        N.check(self._lib.rsb_plan_add_buffer(self._h, channels, scale, C.byref(bid)))
        self.scales[bid.value] = scale
        return Ref(bid.value, 0, channels)

    def conv(self, src: Ref, dst: Ref, weight, bias=None, ln=None, **kw) -> None:
        """Add a conv op.  On the bf16 (tensor-core) plan a conv whose packed kernel would not fit shared memory next
        to the activation stages is split over its output channels into several ops (each re-reads the input; every
        op keeps its slice of bias / PReLU slopes / residual / destination).

        ``ln = (stats, gamma, beta)``: the conv (1 x 1) consumes LayerNorm(src) without that map ever being written —
        ``stats`` is the 8-channel buffer ``layernorm_stats(src, stats)`` filled; gamma goes into the weights and beta into the bias
        here (fp64), the per-pixel mean / rstd are applied in the conv's epilogue (rsb_conv_desc.ln_fold).  A fourth element
        ``eps`` says that ``stats`` holds the raw partial sums a producing conv wrote with ``ln_out=stats`` (ln_fold = 2) instead of
        the statistics op's {rstd, -mean * rstd}.

        ``ln_out=stats`` (bf16 plans, 1 x 1 convs into a plain planar buffer, cout <= 256): the conv also writes the per-pixel
        {sum, sum of squares} of the values it stores into ``stats`` — the LayerNorm statistics of its output without a pass of their
        own (``ln_out_supported``)."""
        if ln is not None:
            stats, gamma, beta = ln[:3]
            if len(ln) > 3:
                kw['ln_raw_eps'] = float(ln[3])
            w64 = (weight.detach().to('cpu', torch.float64).numpy() if isinstance(weight, torch.Tensor) else np.asarray(weight, dtype=np.float64))
            g64 = (gamma.detach().to('cpu', torch.float64).numpy() if isinstance(gamma, torch.Tensor) else np.asarray(gamma, dtype=np.float64))
            b64 = (beta.detach().to('cpu', torch.float64).numpy() if isinstance(beta, torch.Tensor) else np.asarray(beta, dtype=np.float64))
            assert w64.ndim == 4 and w64.shape[2:] == (1, 1), 'a LayerNorm fold needs a 1x1 conv'
            extra = w64[:, :, 0, 0] @ b64
            bias = extra if bias is None else (bias.detach().to('cpu', torch.float64).numpy() if isinstance(bias, torch.Tensor) else np.asarray(bias, dtype=np.float64)) + extra
            weight = w64 * g64.reshape(1, -1, 1, 1)
            kw['ln_stats'] = stats
        w = _f32(weight)
        cout, cin, kh, kw_ = w.shape
        cin16 = (cin + 15) // 16 * 16
        budget = 150 * 1024
        splittable = (self.compute_dtype == torch.bfloat16 and src.buf >= 0 and dst.buf >= 0 and kw.get('dst_ps', 1) == 1
                      and kw.get('dst2') is None and not kw.get('src_upsample2', False) and kw.get('border_bias') is None)
        npad = (cout + 15) // 16 * 16
        if not splittable or (kh * kw_ * cin16 * npad * 2 <= budget and npad <= 256):
            return self._conv_one(src, dst, w, bias, **kw)
        assert kw.get('ln_out') is None, 'a conv that writes LayerNorm sums (ln_out) cannot be split over its output channels'
        per = max(16, min(256, budget // (kh * kw_ * cin16 * 2) // 16 * 16))
        # equal parts (384 -> 192 + 192 rather than 256 + 128): parts of one width over the same source run as ONE N-split launch
        parts = -(-cout // per)
        per = min(per, -(-(-(-cout // parts)) // 16) * 16)
        b = _f32(bias) if bias is not None else None
        slopes = _f32(kw['act_slopes']) if kw.get('act_slopes') is not None else None
        for c0 in range(0, cout, per):
            n = min(per, cout - c0)
            sub = dict(kw)
            if slopes is not None:
                sub['act_slopes'] = slopes[c0:c0 + n]
            for key in ('res1', 'res2'):
                if sub.get(key) is not None:
                    sub[key] = sub[key].slice(c0, n)
            self._conv_one(src, dst.slice(c0, n), w[c0:c0 + n], None if b is None else b[c0:c0 + n], **sub)

    def _conv_one(
        self,
        src: Ref,
        dst: Ref,
        weight,
        bias=None,
        *,
        act: int = N.ACT_NONE,
        act_param: float = 0.0,
        act_slopes=None,
        combine: int = N.COMB_NONE,
        res1: Optional[Ref] = None,
        res2: Optional[Ref] = None,
        alpha: float = 1.0,
        beta1: float = 1.0,
        beta2: float = 1.0,
        in_mean: Sequence[float] = (0.0, 0.0, 0.0, 0.0),
        in_scale: float = 1.0,
        ps: int = 1,
        add_base: bool = False,
        out_scale: float = 1.0,
        out_mean: Sequence[float] = (0.0, 0.0, 0.0, 0.0),
        src_upsample2: bool = False,
        dst_ps: int = 1,
        dst_phase: int = -1,
        dst2: Optional[Ref] = None,
        pad: Optional[tuple] = None,
        border_bias=None,
        ln_stats: Optional[Ref] = None,
        ln_raw_eps: Optional[float] = None,
        ln_out: Optional[Ref] = None,
    ) -> None:
        w = _f32(weight)
        assert w.ndim == 4, 'conv weight must be [cout][cin][kh][kw]'
        cout, cin, kh, kw = w.shape
        b = _f32(bias) if bias is not None else None
        s = _f32(act_slopes) if act_slopes is not None else None
        d = N.ConvDesc()
        d.src_buf, d.src_ch_off, d.cin = src.buf, src.ch_off, cin
        d.dst_buf, d.dst_ch_off, d.cout = dst.buf, dst.ch_off, cout
        if src.buf >= 0:
            assert src.channels == cin, f'conv reads {cin} channels, source range has {src.channels}'
        main_ch = cout // (dst_ps * dst_ps) if (dst_ps > 1 and dst_phase < 0) else (cout - dst2.channels if dst2 is not None else cout)
        if dst.buf >= 0:
            assert dst.channels == main_ch, f'conv writes {main_ch} channels, destination range has {dst.channels}'
        d.kh, d.kw = kh, kw
        d.weight, d.bias = _fptr(w), _fptr(b)
        d.act, d.act_param, d.act_slopes = act, float(act_param), _fptr(s)
        d.combine = combine
        d.res1_buf, d.res1_ch_off = (res1.buf, res1.ch_off) if res1 is not None else (N.NO_BUFFER, 0)
        d.res2_buf, d.res2_ch_off = (res2.buf, res2.ch_off) if res2 is not None else (N.NO_BUFFER, 0)
        d.alpha, d.beta1, d.beta2 = float(alpha), float(beta1), float(beta2)
        mean4 = (list(in_mean) + [0.0] * 4)[:4]
        d.in_mean = (C.c_float * 4)(*mean4)
        d.in_scale = float(in_scale)
        d.ps, d.add_base, d.out_scale = int(ps), int(bool(add_base)), float(out_scale)
        omean4 = (list(out_mean) + [0.0] * 4)[:4]
        d.out_mean = (C.c_float * 4)(*omean4)
        d.src_upsample2 = int(bool(src_upsample2))
        d.dst_ps = int(dst_ps)
        d.dst_phase = int(dst_phase)
        d.pad_t, d.pad_l = (int(pad[0]), int(pad[1])) if pad is not None else (-1, -1)
        bb = _f32(border_bias) if border_bias is not None else None
        if bb is not None:
            assert bb.shape == (16, cout), 'border_bias must be [16][cout]'
        d.border_bias = _fptr(bb)
        d.dst2_buf, d.dst2_ch_off = (dst2.buf, dst2.ch_off) if dst2 is not None else (N.NO_BUFFER, 0)
        d.split_ch = main_ch if dst2 is not None else 0
        d.ln_fold, d.ln_stats_buf = ((2 if ln_raw_eps is not None else 1), ln_stats.buf) if ln_stats is not None else (0, N.NO_BUFFER)
        d.ln_eps = float(ln_raw_eps) if ln_raw_eps is not None else 0.0
        d.ln_out, d.ln_out_buf = (1, ln_out.buf) if ln_out is not None else (0, N.NO_BUFFER)
        N.check(self._lib.rsb_plan_add_conv(self._h, C.byref(d)))

    def groupnorm(self, src: Ref, dst: Ref, groups: int, gamma, beta, eps: float = 1e-5, skip: Optional[Ref] = None) -> None:
        g, b = _f32(gamma), _f32(beta)
        d = N.GroupNormDesc()
        d.src_buf, d.src_ch_off = src.buf, src.ch_off
        d.dst_buf, d.dst_ch_off = dst.buf, dst.ch_off
        d.channels, d.groups, d.eps = src.channels, groups, float(eps)
        d.gamma, d.beta = _fptr(g), _fptr(b)
        d.skip_buf, d.skip_ch_off = (skip.buf, skip.ch_off) if skip is not None else (N.NO_BUFFER, 0)
        N.check(self._lib.rsb_plan_add_groupnorm(self._h, C.byref(d)))

    def op(self, kind: int, src: Ref, dst: Ref, channels: int, *, src2: Optional[Ref] = None, ints: Sequence[int] = (),
           floats: Sequence[float] = (), weights: Sequence = ()) -> None:
        """Token-wise / attention op (rsb_op_desc): LayerNorm, depthwise 3x3, window / channel attention, AIM."""
        d = N.OpDesc()
        d.kind = kind
        d.src_buf, d.src_ch_off = src.buf, src.ch_off
        d.src2_buf, d.src2_ch_off = (src2.buf, src2.ch_off) if src2 is not None else (N.NO_BUFFER, 0)
        d.dst_buf, d.dst_ch_off = dst.buf, dst.ch_off
        d.channels = channels
        for k, v in enumerate(ints):
            d.i[k] = int(v)
        for k, v in enumerate(floats):
            d.f[k] = float(v)
        keep = []
        for k, arr in enumerate(weights):
            a = _f32(arr).reshape(-1)
            keep.append(a)
            d.w[k] = _fptr(a)
            d.wn[k] = a.size
        N.check(self._lib.rsb_plan_add_op(self._h, C.byref(d)))

    def layernorm(self, src: Ref, dst: Ref, gamma, beta, eps: float = 1e-5) -> None:
        self.op(N.OP_LAYERNORM, src, dst, src.channels, floats=(eps,), weights=(gamma, beta))

    def ln_out_supported(self, channels: int) -> bool:
        """Can a 1 x 1 conv producing ``channels`` channels also write their LayerNorm sums (``conv(..., ln_out=stats)``)?"""
        return self.compute_dtype == torch.bfloat16 and channels <= 256

    def layernorm_stats(self, src: Ref, dst: Ref, eps: float = 1e-5) -> None:
        """Per-pixel LayerNorm statistics of ``src`` into the 8-channel buffer ``dst`` (consumed by ``conv(..., ln=(dst, gamma, beta))``)."""
        assert dst.channels == 8 and dst.ch_off == 0, 'the statistics buffer is a whole 8-channel buffer'
        c = src.channels
        self.op(N.OP_LAYERNORM, src, dst, c, ints=(1,), floats=(eps,), weights=(np.ones(c, np.float32), np.zeros(c, np.float32)))

    def rmsnorm_stats(self, src: Ref, dst: Ref, eps: float = 1e-6) -> None:
        """Per-pixel RMSNorm statistics {1 / (|x|_2 / sqrt(C) + eps), 0} of ``src`` into the 8-channel buffer ``dst``: a 1 x 1 conv with
        ``ln=(dst, scale, offset)`` then computes conv(RMSNorm(src)) without the normalised map ever being written."""
        assert dst.channels == 8 and dst.ch_off == 0, 'the statistics buffer is a whole 8-channel buffer'
        c = src.channels
        self.op(N.OP_LAYERNORM, src, dst, c, ints=(2,), floats=(eps,), weights=(np.ones(c, np.float32), np.zeros(c, np.float32)))

    def dwconv3(self, src: Ref, dst: Ref, weight, bias, act: int = N.ACT_NONE, gate: Optional[Ref] = None) -> None:
        self.op(N.OP_DWCONV3, src, dst, src.channels, src2=gate, ints=(act,), weights=(weight, bias))

    def dwconv(self, src: Ref, dst: Ref, weight, bias) -> None:
        """Depthwise K x K conv (K = 3, 5, 7, 9, 11 from the weight's shape [C][1][K][K]; taps that are zero on a whole 8-channel plane are skipped)."""
        k = int(_f32(weight).shape[-1])
        self.op(N.OP_DWCONV3, src, dst, src.channels, ints=(N.ACT_NONE, 0 if k == 3 else k), weights=(weight, bias))

    def rmsnorm(self, src: Ref, dst: Ref, scale, offset, eps: float = 1e-6) -> None:
        self.op(N.OP_RMSNORM, src, dst, src.channels, floats=(eps,), weights=(scale, offset))

    def unshuffle_pool(self, src: Ref, dst: Ref) -> None:
        """dst (5C channels on the half grid) = [PixelUnshuffle(2)(src) | MaxPool2d(2)(src)]."""
        self.op(N.OP_UNSHUFFLE_POOL, src, dst, src.channels)

    def se_shuffle(self, src: Ref, dst: Ref, se=None) -> None:
        """dst (C/4 channels on the twice finer grid) = PixelShuffle(2)(src * gate); ``se`` = (W1, b1, W2, b2) of the squeeze-excitation
        MLP (ReLU, Hardsigmoid) or None for a plain PixelShuffle."""
        if se is None:
            self.op(N.OP_SE_SHUFFLE, src, dst, src.channels, ints=(0,))
        else:
            w1 = _f32(se[0]).reshape(-1, src.channels)
            self.op(N.OP_SE_SHUFFLE, src, dst, src.channels, ints=(w1.shape[0],), weights=(w1, se[1], _f32(se[2]).reshape(src.channels, -1), se[3]))

    def chan_gate(self, src: Ref, dst: Ref, res: Ref, weight, bias, gamma) -> None:
        """dst = src * ((weight . mean_hw(src) + bias) * gamma)[c] + res (GateRV3's simplified channel attention + shortcut)."""
        c = src.channels
        self.op(N.OP_CHAN_GATE, src, dst, c, src2=res, weights=(_f32(weight).reshape(c, c), bias, _f32(gamma).reshape(-1)))

    def chan_affine(self, src: Ref, dst: Ref, scale, res: Optional[Ref] = None) -> None:
        """dst = src * scale[c] (+ res)."""
        self.op(N.OP_CHAN_AFFINE, src, dst, src.channels, src2=res, weights=(_f32(scale).reshape(-1),))

    def finalize(self, device: torch.device) -> 'Plan':
        index = device.index if device.index is not None else torch.cuda.current_device()
        N.check(self._lib.rsb_plan_finalize(self._h, index))
        handle, self._h = self._h, None
        return Plan(handle, self, torch.device('cuda', index))

    def __del__(self):
        if getattr(self, '_h', None):
            self._lib.rsb_plan_destroy(self._h)


class Plan:
    """Finalised native plan bound to one CUDA device."""

    def __init__(self, handle, builder: PlanBuilder, device: torch.device):
        self._lib = builder._lib
        self._h = handle
        self.device = device
        self.compute_dtype = builder.compute_dtype
        self.in_channels, self.out_channels, self.upscale = builder.in_channels, builder.out_channels, builder.upscale
        # workspaces per input shape, least recently used first: edge tiles of tiled_forward alternate between a few shapes and
        # must not re-allocate / re-bind (re-memset, re-encode every tensor map) on every switch
        self._workspaces: 'OrderedDict[tuple, torch.Tensor]' = OrderedDict()
        self._ws_shape = None
        self.force_direct = False
        self._scale_of = dict(builder.scales)
        self._base_divisor = builder.base_divisor
        # CUDA-graph replay of whole forwards into caller-owned tensors: key = every pointer / shape baked into the launches
        self._graphs = {}
        self._graphs_enabled = os.environ.get('RSB_NO_GRAPH') is None
        weakref.finalize(self, self._lib.rsb_plan_destroy, handle)

    @property
    def launches_per_forward(self) -> int:
        return int(self._lib.rsb_plan_launches_per_forward(self._h))

    def flops(self, n: int, h: int, w: int) -> float:
        out = C.c_double()
        N.check(self._lib.rsb_plan_flops(self._h, n, h, w, C.byref(out)))
        return out.value

    def workspace_bytes(self, n: int, h: int, w: int) -> int:
        out = C.c_size_t()
        N.check(self._lib.rsb_plan_workspace_bytes(self._h, n, h, w, C.byref(out)))
        return int(out.value)

    MAX_WORKSPACES = 4

    def _workspace_for(self, n: int, h: int, w: int) -> torch.Tensor:
        key = (n, h, w)
        ws = self._workspaces.get(key)
        if ws is None:
            while len(self._workspaces) >= self.MAX_WORKSPACES:
                _, old = self._workspaces.popitem(last=False)
                # captured graphs that point into the evicted workspace go with it
                self._graphs = {k: g for k, g in self._graphs.items() if k[7] != old.data_ptr()}
                del old
            nbytes = self.workspace_bytes(n, h, w)
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            shift = (-raw.data_ptr()) % 1024
            ws = raw[shift:shift + nbytes]
            self._workspaces[key] = ws
        else:
            self._workspaces.move_to_end(key)
        self._ws_shape = key
        return ws

    @property
    def num_ops(self) -> int:
        return int(self._lib.rsb_plan_num_ops(self._h))

    @property
    def num_direct_convs(self) -> int:
        """bf16 plans: convolutions that fell back to the CUDA-core kernel (0 for every supported configuration)."""
        return int(self._lib.rsb_plan_num_direct_convs(self._h))

    def set_nvtx(self, enable: bool = True) -> None:
        """One NVTX range per op ("rsb op <index> <kernel>") around its launches, for nsys / ncu --nvtx captures."""
        N.check(self._lib.rsb_plan_set_nvtx(self._h, int(bool(enable))))

    def op_info(self, index: int) -> dict:
        """Kernel, launch count, algorithmic FLOPs / HBM bytes of op ``index`` at the shape of the last forward."""
        info = N.OpInfo()
        N.check(self._lib.rsb_plan_op_info(self._h, index, C.byref(info)))
        return dict(kind=info.kind, kernel=(self._lib.rsb_kernel_name(info.kernel) or b'').decode(), fused_next=int(info.fused_next),
                    launches=info.launches, flops=info.flops, bytes=info.bytes)

    @property
    def fused_pairs(self) -> int:
        """Fused conv-pair launches per forward at the shape of the last forward."""
        return sum(1 for i in range(self.num_ops) if self.op_info(i)['fused_next'] and self.op_info(i)['kernel'] == 'conv_pair')

    def launch_units(self):
        """[(op_begin, op_end, info)] — the op ranges that run as one kernel launch group (a fused pair is one unit)."""
        units, i = [], 0
        while i < self.num_ops:
            info = self.op_info(i)
            j = i + 1 + info['fused_next']
            units.append((i, j, info))
            i = j
        return units

    MAX_GRAPHS = 16

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, ops: Optional[tuple] = None, graph: bool = False) -> torch.Tensor:
        """Run the plan (or only ops[0]..ops[1]-1 of it, for per-layer timing) on the current stream.

        ``graph=True`` (needs a caller-owned ``out``): the forward's launches are captured once per (input, output, workspace)
        address combination into a CUDA graph and replayed afterwards — no per-kernel host launch cost (a SPAN 1080p frame is
        23 launches, DAT 675).  Streaming callers that rotate over a few device slots (FramePipeline, bench.py) hit the cache
        after their first pass; anything that fails to capture falls back to plain launches for good."""
        if x.device != self.device:
            raise RuntimeError(f'plan lives on {self.device}, input is on {x.device}')
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f'expected NCHW input with {self.in_channels} channels, got {tuple(x.shape)}')
        if x.dtype not in _TORCH_TO_RSB:
            raise RuntimeError(f'unsupported input dtype {x.dtype}')
        stable_x = x.is_contiguous()  # a temporary contiguous copy has no stable address worth a graph
        x = x.contiguous()
        n, _, h, w = x.shape
        shape = (n, self.out_channels, h * self.upscale, w * self.upscale)
        caller_out = out is not None
        if out is None:
            out = torch.empty(shape, dtype=x.dtype, device=x.device)
        elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype not in _TORCH_TO_RSB:
            raise RuntimeError('out tensor has the wrong shape/layout')
        ws = self._workspace_for(n, h, w)
        begin, end = ops if ops is not None else (0, self.num_ops)

        def launch():
            stream = torch.cuda.current_stream(self.device).cuda_stream
            N.check(
                self._lib.rsb_plan_forward_ops(
                    self._h, x.data_ptr(), _TORCH_TO_RSB[x.dtype], n, h, w, out.data_ptr(), _TORCH_TO_RSB[out.dtype],
                    ws.data_ptr(), ws.numel(), stream, int(self.force_direct), begin, end,
                )
            )

        if graph and caller_out and stable_x and ops is None and self._graphs_enabled and not torch.cuda.is_current_stream_capturing():
            key = (x.data_ptr(), out.data_ptr(), x.dtype, out.dtype, n, h, w, ws.data_ptr(), int(self.force_direct))
            g = self._graphs.get(key)
            if g is None and len(self._graphs) < self.MAX_GRAPHS:
                launch()  # binds workspace / tensor maps for this shape outside the capture (and produces this call's result)
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode='thread_local'):
                        launch()
                except N.NativeError:
                    raise  # a real engine error (bad workspace, failed launch) is never masked by the graph fallback
                except RuntimeError as exc:
                    # capture itself failed (e.g. an allocation inside the capture): graphs are an optimisation, fall back to
                    # plain launches for this plan — and say so, once
                    self._graphs_enabled = False
                    self._graphs.clear()
                    _log.warning('resselt_b200: CUDA-graph capture failed (%s); this plan falls back to plain launches', exc)
                    torch.cuda.synchronize(self.device)
                    return out  # the eager launch above already produced the result
                self._graphs[key] = g
                return out
            if g is not None:
                g.replay()
                return out
        launch()
        return out

    def capture(self, x: torch.Tensor, out: torch.Tensor) -> None:
        """Warm-up API for streaming callers: run one forward into ``out`` and capture the CUDA graph for this (x, out) pair now —
        outside the streaming loop, where the capture's device synchronisation would stall the copy streams."""
        self.forward(x, out=out, graph=True)

    def read_buffer(self, ref: Ref) -> torch.Tensor:
        """Debug/test helper: fp32 NCHW copy of a buffer range after the last forward."""
        n, h, w = self._ws_shape
        scale = self._scale_of[ref.buf]
        dst = torch.empty((n, ref.channels, h // self._base_divisor * scale, w // self._base_divisor * scale), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self._lib.rsb_plan_read_buffer(self._h, ref.buf, ref.ch_off, ref.channels, dst.data_ptr(), stream))
        return dst
