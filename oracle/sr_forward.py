"""Functional CPU restatements of the reference forwards (test oracle — see package docstring).

Every function takes ``sd`` (a state dict with the reference's key names), an NCHW tensor ``x`` and
a compute ``dtype`` (torch.float32 to mirror the reference's fp32 forward, torch.float64 as the
high-precision arbiter).  Hyper-parameters are re-derived from tensor shapes, like the reference's
loaders do.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Mapping

import torch
import torch.nn.functional as F

SD = Mapping[str, torch.Tensor]


def _seq_len(sd: SD, key: str) -> int:
    idx = [int(k[len(key) + 1:].split('.', 1)[0]) for k in sd if k.startswith(key + '.')]
    return max(idx) + 1 if idx else 0


def _conv(sd: SD, name: str, x: torch.Tensor, pad: int) -> torch.Tensor:
    return F.conv2d(x, sd[name + '.weight'].to(x.dtype), sd[name + '.bias'].to(x.dtype), padding=pad)


def conv3xc_merged(sd: SD, prefix: str, dtype: torch.dtype):
    """Equivalent 3x3 kernel/bias of a Conv3XC in eval mode.

    Follows Conv3XC.update_params (/root/reference/resselt/archs/span/arch.py:124-150, identical in
    /root/reference/resselt/archs/spanplus/arch.py:66-92): 1x1(w1,b1) -> 3x3(w2,b2) -> 1x1(w3,b3)
    composed, plus the 1x1 skip ``sk`` added at the centre tap.  ``eval_conv.*`` from the checkpoint
    is ignored because the reference overwrites it in every forward (:152-154).
    """
    g = lambda k: sd[f'{prefix}.{k}'].to(dtype)
    w1, b1 = g('conv.0.weight')[:, :, 0, 0], g('conv.0.bias')
    w2, b2 = g('conv.1.weight'), g('conv.1.bias')
    w3, b3 = g('conv.2.weight')[:, :, 0, 0], g('conv.2.bias')
    k = torch.einsum('on,nmhw,mi->oihw', w3, w2, w1)
    k[:, :, 1, 1] += g('sk.weight')[:, :, 0, 0]
    b = w3 @ ((w2 * b1.view(1, -1, 1, 1)).sum((1, 2, 3)) + b2) + b3 + g('sk.bias')
    return k, b


def _conv3xc(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    k, b = conv3xc_merged(sd, prefix, x.dtype)
    return F.conv2d(x, k, b, padding=1)


def _spab(sd: SD, prefix: str, x: torch.Tensor, act: Callable):
    """SPAB.forward (/root/reference/resselt/archs/span/arch.py:167-180; spanplus/arch.py:117-130).
    Returns (gated output, *activated* first conv output) — the activation is in-place in the reference."""
    o1 = act(_conv3xc(sd, f'{prefix}.c1_r', x))
    o2 = act(_conv3xc(sd, f'{prefix}.c2_r', o1))
    o3 = _conv3xc(sd, f'{prefix}.c3_r', o2)
    return (o3 + x) * (torch.sigmoid(o3) - 0.5), o1


def span_forward(sd: SD, x: torch.Tensor, dtype=torch.float32, img_range: float = 255.0,
                 rgb_mean=(0.4488, 0.4371, 0.4040)) -> torch.Tensor:
    """SPAN.forward (/root/reference/resselt/archs/span/arch.py:231-250); loader defaults for
    img_range / rgb_mean from /root/reference/resselt/archs/span/__init__.py:27-29."""
    x = x.to(dtype)
    if 'no_norm' not in sd:
        x = (x - torch.tensor(rgb_mean, dtype=dtype).view(1, 3, 1, 1)) * img_range
    feat = _conv3xc(sd, 'conv_1', x)
    b1, _ = _spab(sd, 'block_1', feat, F.silu)
    b = b1
    for i in (2, 3, 4, 5):
        b, _ = _spab(sd, f'block_{i}', b, F.silu)
    b6, o1 = _spab(sd, 'block_6', b, F.silu)
    tail = _conv3xc(sd, 'conv_2', b6)
    out = _conv(sd, 'conv_cat', torch.cat([feat, tail, b1, o1], 1), 0)
    in_ch = sd['conv_1.sk.weight'].shape[1]
    r = math.isqrt(sd['upsampler.0.weight'].shape[0] // in_ch)
    return F.pixel_shuffle(_conv(sd, 'upsampler.0', out, 1), r)


def spanplus_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """SpanPlus.forward with the 'ps' upsampler, eval mode
    (/root/reference/resselt/archs/spanplus/arch.py:199-201, SPABS.forward :146-151)."""
    x = x.to(dtype)
    groups = _seq_len(sd, 'feats') - 1
    cur = _conv3xc(sd, 'feats.0', x)
    for g in range(1, groups + 1):
        pre = f'feats.{g}'
        b1, _ = _spab(sd, f'{pre}.block_1', cur, F.mish)
        t = b1
        for i in range(_seq_len(sd, f'{pre}.block_n')):
            t, _ = _spab(sd, f'{pre}.block_n.{i}', t, F.mish)
        end, o1 = _spab(sd, f'{pre}.block_end', t, F.mish)
        tail = _conv3xc(sd, f'{pre}.conv_2', end)
        cur = _conv(sd, f'{pre}.conv_cat', torch.cat([cur, tail, b1, o1], 1), 0)
    in_ch = sd['feats.0.eval_conv.weight'].shape[1]
    r = math.isqrt(sd['upsampler.0.weight'].shape[0] // in_ch)
    return F.pixel_shuffle(_conv(sd, 'upsampler.0', cur, 1), r)


def compact_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """SRVGGNetCompact.forward (/root/reference/resselt/archs/compact/arch.py:56-65)."""
    x = x.to(dtype)
    top = _seq_len(sd, 'body') - 1
    out = x
    for i in range(0, top, 2):
        out = F.prelu(_conv(sd, f'body.{i}', out, 1), sd[f'body.{i + 1}.weight'].to(dtype))
    out = _conv(sd, f'body.{top}', out, 1)
    r = math.isqrt(out.shape[1] // x.shape[1])
    return F.pixel_shuffle(out, r) + F.interpolate(x, scale_factor=r, mode='nearest')


_FORWARDS: Dict[str, Callable] = {
    'SPAN': span_forward,
    'SPANPlus': spanplus_forward,
    'Compact': compact_forward,
}


def forward_by_name(name: str, sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    with torch.inference_mode():
        return _FORWARDS[name](sd, x, dtype)
