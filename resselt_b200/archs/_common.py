"""Helpers shared by the architecture plugins: parameter-spec builders and weight algebra."""
from __future__ import annotations

from typing import Dict, List

import torch

from ..engine import ParamSpec


def conv_specs(prefix: str, cin: int, cout: int, k: int, gain: float = 1.0) -> List[ParamSpec]:
    """Names/shapes of one ``nn.Conv2d(cin, cout, k)`` with bias (``gain`` scales the random-init weight range)."""
    fan_in = cin * k * k
    kind = 'conv_w' if gain == 1.0 else f'conv_w*{gain}'
    return [(f'{prefix}.weight', (cout, cin, k, k), kind), (f'{prefix}.bias', (cout,), f'bias:{fan_in}')]


def conv3xc_specs(prefix: str, cin: int, cout: int, gain: int = 2) -> List[ParamSpec]:
    """Parameter names of a re-parameterisable Conv3XC (reference span/arch.py:59-122):
    ``sk`` 1x1, ``conv.0`` 1x1 expand, ``conv.1`` 3x3, ``conv.2`` 1x1 reduce, and the (dead) ``eval_conv``."""
    return (
        conv_specs(f'{prefix}.sk', cin, cout, 1)
        + conv_specs(f'{prefix}.conv.0', cin, cin * gain, 1)
        + conv_specs(f'{prefix}.conv.1', cin * gain, cout * gain, 3)
        + conv_specs(f'{prefix}.conv.2', cout * gain, cout, 1)
        + conv_specs(f'{prefix}.eval_conv', cin, cout, 3)
    )


def merge_conv3xc(w: Dict[str, torch.Tensor], prefix: str):
    """Collapse 1x1 -> 3x3 -> 1x1 plus the parallel 1x1 skip into a single 3x3 (weight, bias), in fp64.

    Closed form of what the reference recomputes on every forward (span/arch.py:124-150):
        W[o,i,:,:] = sum_{n,m} w3[o,n] * w2[n,m,:,:] * w1[m,i]   (+ sk_w[o,i] at the centre tap)
        b[o]       = sum_n w3[o,n] * (sum_{m,kh,kw} w2[n,m,kh,kw] * b1[m] + b2[n]) + b3[o] + sk_b[o]
    The checkpoint's ``eval_conv.*`` tensors are never read: the reference overwrites them before use.
    """
    f64 = torch.float64
    w1, b1 = w[f'{prefix}.conv.0.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.conv.0.bias'].to(f64)
    w2, b2 = w[f'{prefix}.conv.1.weight'].to(f64), w[f'{prefix}.conv.1.bias'].to(f64)
    w3, b3 = w[f'{prefix}.conv.2.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.conv.2.bias'].to(f64)
    sk_w, sk_b = w[f'{prefix}.sk.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.sk.bias'].to(f64)
    merged = torch.einsum('on,nmhw,mi->oihw', w3, w2, w1)
    merged[:, :, 1, 1] += sk_w
    inner_bias = torch.einsum('nmhw,m->n', w2, b1) + b2
    bias = w3 @ inner_bias + b3 + sk_b
    return merged, bias


def merge_pointwise_into_conv(w1: torch.Tensor, b1, wk: torch.Tensor, bk):
    """Collapse ``conv_kxk(conv_1x1(x))`` (no activation in between, zero padding on the k x k conv) into ONE k x k conv, fp64.

    Returns (weight [O][C][k][k], bias [O], border_bias [16][O]).  W[o,c,ky,kx] = sum_m wk[o,m,ky,kx] w1[m,c]; the 1x1 conv's bias
    reaches an output pixel through every tap that lies INSIDE the image (the k x k conv zero-pads the 1x1 conv's output, bias
    included), so bias[o] = bk[o] + sum_{m,taps} wk[o,m,tap] b1[m] holds for interior pixels and border pixels get
    border_bias[mask][o] = - sum_{m, taps outside under mask} wk[o,m,tap] b1[m], mask = top | bottom << 1 | left << 2 | right << 3
    (rsb_conv_desc.border_bias).  Used for SPAN's / SPANPlus' / SpanPP's ``upsampler(conv_cat(cat))`` (span/arch.py:247-248):
    one pass over the 192-channel concat instead of a 1x1 pass plus a 3x3 pass."""
    f64 = torch.float64
    w1, wk = w1.to(f64)[:, :, 0, 0], wk.to(f64)
    b1 = torch.zeros(w1.shape[0], dtype=f64) if b1 is None else b1.to(f64)
    bk = torch.zeros(wk.shape[0], dtype=f64) if bk is None else bk.to(f64)
    k = wk.shape[-1]
    r = k // 2
    weight = torch.einsum('omyx,mc->ocyx', wk, w1)
    per_tap = torch.einsum('omyx,m->oyx', wk, b1)  # what each tap contributes of the 1x1 conv's bias
    bias = bk + per_tap.sum((1, 2))
    border = torch.zeros(16, wk.shape[0], dtype=f64)
    for mask in range(1, 16):
        top, bottom, left, right = mask & 1, mask & 2, mask & 4, mask & 8
        for ky in range(k):
            for kx in range(k):
                dy, dx = ky - r, kx - r
                # a k x k conv with k > 3 would need distance-dependent classes; this table is exact for k == 3 (and k == 1)
                if (dy < 0 and top) or (dy > 0 and bottom) or (dx < 0 and left) or (dx > 0 and right):
                    border[mask] -= per_tap[:, ky, kx]
    return weight, bias, border


# ------------------------------------------------------------------------------------------------ DySample head
HEAD_PAD = 32  # channel stride of a head in the head-padded q / k / v layout (csrc/winattn_tc.cu)


def winattn_head_padded(compute_dtype, dim: int, heads: int, split) -> bool:
    """True when window attention can run on the tcgen05 kernel (rsb_op_desc.i[5] = 32): bf16 plan, even heads, head_dim < 32,
    windows of 64 / 128 / 256 tokens whose sides are multiples of 8 (mirrors ``winattn_tc_supported`` in csrc/winattn_tc.cu)."""
    hs, ws = int(split[0]), int(split[1])
    return (compute_dtype == torch.bfloat16 and heads % 2 == 0 and dim % heads == 0 and dim // heads < HEAD_PAD
            and hs * ws in (64, 128, 256) and hs % 8 == 0 and ws % 8 == 0 and max(hs, ws) <= 32)


def head_pad_index(dim: int, heads: int) -> torch.Tensor:
    """Position of channel c (head c // d, dim c % d) in the head-padded layout: (c // d) * 32 + c % d."""
    d = dim // heads
    c = torch.arange(dim)
    return (c // d) * HEAD_PAD + c % d


def pad_head_rows(t, dim: int, heads: int):
    """Scatter dim 0 of a [dim, ...] tensor (an output-channel axis: weight rows, bias) into [heads * 32, ...], zeros elsewhere."""
    if t is None:
        return None
    out = t.new_zeros((heads * HEAD_PAD,) + tuple(t.shape[1:]))
    out[head_pad_index(dim, heads)] = t
    return out


def pad_head_cols(t, dim: int, heads: int):
    """Scatter dim 1 of a [out, dim, ...] tensor (an input-channel axis) into [out, heads * 32, ...], zeros elsewhere."""
    out = t.new_zeros((t.shape[0], heads * HEAD_PAD) + tuple(t.shape[2:]))
    out[:, head_pad_index(dim, heads)] = t
    return out


def dysample_init_pos(scale: int, groups: int) -> torch.Tensor:
    """The ``init_pos`` buffer exactly as DySample._init_pos builds it (/root/reference/resselt/utilities/dysample.py:42-44)."""
    h = torch.arange((-scale + 1) / 2, (scale - 1) / 2 + 1) / scale
    return torch.stack(torch.meshgrid([h, h], indexing='ij')).transpose(1, 2).repeat(1, groups, 1).reshape(1, -1, 1, 1)


def dysample_specs(prefix: str, in_channels: int, out_ch: int, scale: int, groups: int = 4, end_convolution: bool = True) -> List[ParamSpec]:
    """Parameter / buffer names of a DySample module (dysample.py:12-40): end_conv (1x1), offset (1x1), scope (1x1, no bias), init_pos."""
    if in_channels < groups or in_channels % groups != 0:
        raise ValueError('Incorrect in_channels and groups values.')  # dysample.py:23-27
    k = 2 * groups * scale * scale
    specs: List[ParamSpec] = []
    if end_convolution:
        specs += conv_specs(f'{prefix}.end_conv', in_channels, out_ch, 1, gain=3.0)
    specs += [(f'{prefix}.offset.weight', (k, in_channels, 1, 1), 'conv_w*2.0'), (f'{prefix}.offset.bias', (k,), f'bias:{in_channels}')]
    specs += [(f'{prefix}.scope.weight', (k, in_channels, 1, 1), 'conv_w*2.0')]
    specs += [(f'{prefix}.init_pos', dysample_init_pos(scale, groups), 'buffer_tensor')]
    return specs


def emit_dysample(pb, w: Dict[str, torch.Tensor], prefix: str, x, out_ch: int, scale: int, groups: int = 4) -> None:
    """DySample.forward (dysample.py:46-83) on the engine: ``0.5 * offset(x)`` and ``scope(x)`` come out of ONE 1x1 conv op (the 0.5
    folded into the weights), the per-group end_conv projections out of another; sigmoid gate, sampling and the sum over groups
    are one fused op writing the caller's NCHW output.  ``x`` is the feature buffer range the head reads (it must hold exactly the head's input channels)."""
    from ..engine import OUTPUT
    from ..engine import native as N

    k = 2 * groups * scale * scale
    kp = (k + 7) // 8 * 8  # scope starts on a plane boundary
    # one conv op produces [0.5 * offset | scope]; the head applies sigmoid(scope) to the offsets it reads
    ow, ob = w[f'{prefix}.offset.weight'], w[f'{prefix}.offset.bias']
    both_w = torch.zeros(kp + k, x.channels, 1, 1, dtype=ow.dtype)
    both_b = torch.zeros(kp + k, dtype=ow.dtype)
    both_w[:k], both_b[:k] = 0.5 * ow, 0.5 * ob
    both_w[kp:] = w[f'{prefix}.scope.weight']
    off = pb.buffer(kp + k, scale=pb.scales[x.buf])
    pb.conv(x, off, both_w, both_b)
    # sampling (bilinear, per group) and the 1x1 end_conv are both linear and commute: project every group through its slice of
    # end_conv on the low-res grid first (4 channels per group, out_ch of them used), then gather 4 values per neighbour instead
    # of the group's feature channels
    end_w = w[f'{prefix}.end_conv.weight'].reshape(out_ch, x.channels)
    cg = x.channels // groups
    zw = torch.zeros(4 * groups, x.channels, 1, 1, dtype=end_w.dtype)
    for g in range(groups):
        zw[4 * g:4 * g + out_ch, g * cg:(g + 1) * cg, 0, 0] = end_w[:, g * cg:(g + 1) * cg]
    z = pb.buffer(4 * groups, scale=pb.scales[x.buf])
    pb.conv(x, z, zw, None)
    pb.op(N.OP_DYSAMPLE, z, OUTPUT, 4 * groups, src2=off, ints=(groups, scale, out_ch, 1, kp),
          weights=(w[f'{prefix}.init_pos'], end_w, w[f'{prefix}.end_conv.bias']))


# ------------------------------------------------------------------------------------------------ '1conv' / '3conv' residual connections
def resi_conv_specs(prefix: str, dim: int, resi_connection: str) -> List[ParamSpec]:
    """The conv closing a residual group of SwinIR / DAT: one 3x3 ('1conv') or 3x3 -> lrelu(0.2) -> 1x1 -> lrelu(0.2) -> 3x3 through a
    dim/4 bottleneck ('3conv') (/root/reference/resselt/archs/swinir/arch.py:564-574, dat/arch.py:750-759)."""
    if resi_connection == '1conv':
        return conv_specs(prefix, dim, dim, 3)
    # random init only: the bottleneck attenuates the signal, a gain keeps the output range of seeded test models sane
    return (conv_specs(f'{prefix}.0', dim, dim // 4, 3, gain=2.0) + conv_specs(f'{prefix}.2', dim // 4, dim // 4, 1, gain=2.0)
            + conv_specs(f'{prefix}.4', dim // 4, dim, 3, gain=2.0))


def emit_resi_conv(pb, w: Dict[str, torch.Tensor], name: str, resi_connection: str, src, dst, res, tmp_a, tmp_b) -> None:
    """dst = conv(src) + res with conv = '1conv' or '3conv'; tmp_a / tmp_b: dim/4-channel scratch buffers for '3conv'."""
    from ..engine import native as N

    if resi_connection == '1conv':
        pb.conv(src, dst, w[f'{name}.weight'], w[f'{name}.bias'], combine=N.COMB_AXPY, res1=res)
        return
    lrelu = dict(act=N.ACT_LRELU, act_param=0.2)
    pb.conv(src, tmp_a, w[f'{name}.0.weight'], w[f'{name}.0.bias'], **lrelu)
    pb.conv(tmp_a, tmp_b, w[f'{name}.2.weight'], w[f'{name}.2.bias'], **lrelu)
    pb.conv(tmp_b, dst, w[f'{name}.4.weight'], w[f'{name}.4.bias'], combine=N.COMB_AXPY, res1=res)
