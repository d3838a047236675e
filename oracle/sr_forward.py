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


def _esrgan_keys(sd: SD):
    if 'model.0.weight' in sd:
        nb = _seq_len(sd, 'model.1.sub') - 1
        n_up = (_seq_len(sd, 'model') - 5) // 3
        return dict(first='model.0', rdb=lambda i, j, k: f'model.1.sub.{i}.RDB{j}.conv{k}.0', one=lambda i, j: f'model.1.sub.{i}.RDB{j}.conv1x1',
                    trunk=f'model.1.sub.{nb}', up=lambda u: f'model.{3 * u}', hr=f'model.{3 * n_up + 2}', last=f'model.{3 * n_up + 4}'), nb, n_up
    new = 'conv_body.weight' in sd
    body = 'body' if new else 'RRDB_trunk'
    nb = _seq_len(sd, body)
    up = 'conv_up' if new else 'upconv'
    n_up = sum(1 for u in range(1, 6) if f'{up}{u}.weight' in sd)
    rdb = (lambda i, j, k: f'body.{i}.rdb{j}.conv{k}') if new else (lambda i, j, k: f'RRDB_trunk.{i}.RDB{j}.conv{k}')
    return dict(first='conv_first', rdb=rdb, one=None, trunk='conv_body' if new else 'trunk_conv', up=lambda u: f'{up}{u}',
                hr='conv_hr' if new else 'HRconv', last='conv_last'), nb, n_up


def esrgan_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """RRDBNet.forward (/root/reference/resselt/archs/esrgan/arch.py:129-138) over the flattened Sequential of
    :72-127; RRDB / ResidualDenseBlock_5C per /root/reference/resselt/utilities/block.py:340-344, :454-465 (incl. the
    ESRGAN+ conv1x1 paths :457-463); upconv_block = nearest x2 -> conv -> LeakyReLU(0.2) (:510-537)."""
    x = x.to(dtype)
    K, nb, n_up = _esrgan_keys(sd)
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    in_nc, out_nc = sd[K['first'] + '.weight'].shape[1], sd[K['last'] + '.weight'].shape[0]
    factor = int(math.sqrt(in_nc / out_nc)) if in_nc in (out_nc * 4, out_nc * 16) else None
    h, w = x.shape[-2:]
    if factor:  # Real-ESRGAN x2/x1: reflect-pad, pixel-unshuffle (:130-134)
        x = F.pad(x, (0, (factor - w % factor) % factor, 0, (factor - h % factor) % factor), 'reflect')
        x = F.pixel_unshuffle(x, factor)
    plus = any('.conv1x1.' in k for k in sd)
    fea = _conv(sd, K['first'], x, 1)
    t = fea
    for i in range(nb):
        rrdb_in = t
        for j in (1, 2, 3):
            c = lambda k, inp: _conv(sd, K['rdb'](i, j, k), inp, 1)
            x0 = t
            x1 = lrelu(c(1, x0))
            x2 = lrelu(c(2, torch.cat((x0, x1), 1)))
            if plus:
                x2 = x2 + F.conv2d(x0, sd[K['one'](i, j) + '.weight'].to(dtype))
            x3 = lrelu(c(3, torch.cat((x0, x1, x2), 1)))
            x4 = lrelu(c(4, torch.cat((x0, x1, x2, x3), 1)))
            if plus:
                x4 = x4 + x2
            x5 = c(5, torch.cat((x0, x1, x2, x3, x4), 1))
            t = x5 * 0.2 + x0
        t = t * 0.2 + rrdb_in
    t = fea + _conv(sd, K['trunk'], t, 1)
    for u in range(1, n_up + 1):
        t = lrelu(_conv(sd, K['up'](u), F.interpolate(t, scale_factor=2, mode='nearest'), 1))
    t = lrelu(_conv(sd, K['hr'], t, 1))
    y = _conv(sd, K['last'], t, 1)
    if factor:
        s = (2 ** n_up) // factor
        y = y[:, :, : h * s, : w * s]
    return y


def realplksr_forward(sd: SD, x: torch.Tensor, dtype=torch.float32, norm_groups: int = 4) -> torch.Tensor:
    """realplksr.forward, eval mode, PixelShuffle head (/root/reference/resselt/archs/plksr/rplksr.py:145-147);
    PLKBlock.forward :85-93 = DCCM (:10-19) -> in-place 17x17 conv on the first pdim channels (:36-37) ->
    EA x*sigmoid(conv(x)) (:48-49) -> 1x1 refine -> GroupNorm(4) -> + skip.  norm_groups is fixed to 4 by the loader
    (/root/reference/resselt/archs/plksr/__init__.py:116)."""
    x = x.to(dtype)
    total = _seq_len(sd, 'feats')
    t = _conv(sd, 'feats.0', x, 1)
    for i in range(1, total - 2):
        p = f'feats.{i}'
        skip = t
        t = _conv(sd, f'{p}.channel_mixer.2', F.mish(_conv(sd, f'{p}.channel_mixer.0', t, 1)), 1)
        lk_w = sd[f'{p}.lk.conv.weight']
        pdim, k = lk_w.shape[0], lk_w.shape[2]
        t = torch.cat([_conv(sd, f'{p}.lk.conv', t[:, :pdim], k // 2), t[:, pdim:]], 1)
        if f'{p}.attn.f.0.weight' in sd:
            t = t * torch.sigmoid(_conv(sd, f'{p}.attn.f.0', t, 1))
        t = _conv(sd, f'{p}.refine', t, 0)
        t = F.group_norm(t, norm_groups, sd[f'{p}.norm.weight'].to(dtype), sd[f'{p}.norm.bias'].to(dtype), 1e-5)
        t = t + skip
    t = _conv(sd, f'feats.{total - 1}', t, 1)
    r2 = t.shape[1] // x.shape[1]
    return F.pixel_shuffle(t + torch.repeat_interleave(x, r2, dim=1), math.isqrt(r2))


_FORWARDS: Dict[str, Callable] = {
    'RealPLKSR': realplksr_forward,
    'ESRGAN': esrgan_forward,
    'SPAN': span_forward,
    'SPANPlus': spanplus_forward,
    'Compact': compact_forward,
}


def forward_by_name(name: str, sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    with torch.inference_mode():
        return _FORWARDS[name](sd, x, dtype)
