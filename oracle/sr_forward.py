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


def repconv_merged(sd: SD, prefix: str, dtype: torch.dtype):
    """Fused 3x3 kernel / bias of a RepConv in eval mode: RepConv.fuse (/root/reference/resselt/archs/spanpp/arch.py:164-173) =
    alpha[0] * SeqConv3x3.rep_params (:136-149) + alpha[1] * conv2 + alpha[2] * Conv3XC.update_params (:61-97)."""
    g = lambda k: sd[f'{prefix}.{k}'].to(dtype)
    k0, k1 = g('conv1.k0'), g('conv1.k1')
    rk = F.conv2d(k1, k0.permute(1, 0, 2, 3))
    rb = F.conv2d(torch.ones(1, k0.shape[0], 3, 3, dtype=dtype) * g('conv1.b0').view(1, -1, 1, 1), k1).view(-1) + g('conv1.b1')
    k3, b3 = conv3xc_merged(sd, f'{prefix}.conv3', dtype)
    a = g('alpha')
    return a[0] * rk + a[1] * g('conv2.weight') + a[2] * k3, a[0] * rb + a[1] * g('conv2.bias') + a[2] * b3


def _igconv_kernel(sd: SD, p: str, dim: int, ksize: int, scale: int, max_s: int, dtype) -> torch.Tensor:
    """IGConv._implicit_representation_latent (/root/reference/resselt/archs/spanpp/arch.py:289-312) with make_coord (:219-231)."""
    g = lambda k: sd[f'{p}.{k}'].to(dtype)
    seq = -1 + 1 / scale + (2 / scale) * torch.arange(scale, dtype=dtype)
    coords = torch.stack(torch.meshgrid(seq, seq, indexing='ij'), dim=-1).flip(-1).permute(2, 0, 1).unsqueeze(0)  # [1, (x, y), s, s]
    r = torch.ones(1, 1, scale, scale, dtype=dtype) / min(scale, max_s) * 2
    freq = g('freq').repeat(1, 1, scale, scale)
    f1, f2 = freq.chunk(2, dim=1)
    freq = f1 * coords[:, :1] + f2 * coords[:, 1:] + F.conv2d(r, g('phase.weight'), g('phase.bias'))
    t = torch.cat([torch.cos(torch.pi * freq), torch.sin(torch.pi * freq)], dim=1) * g('amplitude').repeat(1, 1, scale, scale)
    n = _seq_len(sd, f'{p}.query_kernel')
    for i in range(0, n, 2):
        t = F.conv2d(t, g(f'query_kernel.{i}.weight'), g(f'query_kernel.{i}.bias'))
        t = F.relu(t) if i + 1 < n else t
    k, rgb = t.shape[0], t.shape[1]
    return t.view(dim, ksize, ksize, rgb, scale, scale).permute(3, 4, 5, 0, 1, 2).reshape(rgb * scale * scale, dim, ksize, ksize)


def spanpp_forward(sd: SD, x: torch.Tensor, dtype=torch.float32, scale: int = 2) -> torch.Tensor:
    """SpanPP.forward in eval mode (/root/reference/resselt/archs/spanpp/arch.py:358-373; SPAB :195-216; IGConv.forward :289-297)."""
    x = x.to(dtype)

    def rep(prefix, t):
        k, b = repconv_merged(sd, prefix, dtype)
        return F.conv2d(t, k, b, padding=1)

    def spab(prefix, t):
        o1 = F.silu(rep(f'{prefix}.c1_r', t))          # SiLU(inplace=True): the block's second result is the activated tensor
        o3 = rep(f'{prefix}.c3_r', F.silu(rep(f'{prefix}.c2_r', o1)))
        return (o3 + t) * (torch.sigmoid(o3) - 0.5), o1

    feat = rep('conv0', x)
    b1, _ = spab('block_1', feat)
    t = b1
    for i in (2, 3, 4, 5):
        t, _ = spab(f'block_{i}', t)
    b6, o1 = spab('block_6', t)
    out = _conv(sd, 'conv_cat', torch.cat([feat, rep('conv_2', b6), b1, o1], 1), 0)
    dim = feat.shape[1]
    scales = sd['MetaIGConv'].tolist() if 'MetaIGConv' in sd else [1, 2, 3, 4]
    kernel = _igconv_kernel(sd, 'upsampler', dim, 3, scale, max(scales), dtype)
    return F.pixel_shuffle(F.conv2d(out, kernel, None, padding=1), scale)


def rtmosr_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """RTMoSR.forward in eval mode (/root/reference/resselt/archs/rtmosr/arch.py:375-387; GatedCNNBlock :321-326, ParPixelUnshuffle :284-292,
    OmniShift.reparam_5x5 :249-269, CSELayer :18-21, RMSNorm :32-38, RepConv.fuse :171-180)."""
    x = x.to(dtype)
    g = lambda k: sd[k].to(dtype)
    unshuffle = 0
    if 'to_feat.1.alpha' in sd:
        unshuffle = math.isqrt(sd['to_feat.1.conv_3x3_rep.weight'].shape[1] // 3)
        scale, inner, feat = 4 // unshuffle, 4, 'to_feat.1'
    else:
        scale = inner = math.isqrt(sd['to_img.0.conv_3x3_rep.weight'].shape[0] // 3)
        feat = 'to_feat'
    pad = 2 * max(unshuffle, 1)
    b, _, h, w = x.shape

    def rep(prefix, t):
        k, bias = repconv_merged(sd, prefix, dtype)
        return F.conv2d(t, k, bias, padding=1)

    out = F.pad(x, (0, (pad - w % pad) % pad, 0, (pad - h % pad) % pad), 'reflect')
    out = rep(feat, F.pixel_unshuffle(out, unshuffle) if unshuffle else out)
    dim = out.shape[1]
    for i in range(_seq_len(sd, 'body')):
        p = f'body.{i}'
        short = out
        rms = out.norm(2, dim=1, keepdim=True) * dim ** -0.5
        t = g(f'{p}.norm.scale')[:, None, None] * (out / (rms + 1e-6)) + g(f'{p}.norm.offset')[:, None, None]
        t = rep(f'{p}.fc1', t)
        hidden = t.shape[1] // 2
        gate, ident, c = torch.split(t, [hidden, hidden - dim, dim], dim=1)
        c = F.pixel_unshuffle(c, 2) + rep(f'{p}.conv.0.poll.1', F.max_pool2d(c, 2, 2))
        q = f'{p}.conv.1'
        a = [g(f'{q}.alpha{k}').reshape(-1, 1, 1, 1) for k in (1, 2, 3, 4)]
        k5 = (a[0] * F.pad(torch.ones_like(g(f'{q}.conv1x1.weight')), (2, 2, 2, 2)) + a[1] * F.pad(g(f'{q}.conv1x1.weight'), (2, 2, 2, 2))
              + a[2] * F.pad(g(f'{q}.conv3x3.weight'), (1, 1, 1, 1)) + a[3] * g(f'{q}.conv5x5.weight'))
        b5 = a[1].flatten() * g(f'{q}.conv1x1.bias') + a[2].flatten() * g(f'{q}.conv3x3.bias') + a[3].flatten() * g(f'{q}.conv5x5.bias')
        c = F.conv2d(c, k5, b5, padding=2, groups=c.shape[1])
        if f'{p}.conv.2.squeezing.0.weight' in sd:
            sq = c.mean(dim=(2, 3), keepdim=True)
            sq = F.hardsigmoid(_conv(sd, f'{p}.conv.2.squeezing.2', F.relu(_conv(sd, f'{p}.conv.2.squeezing.0', sq, 0)), 0))
            c = c * sq
        c = F.pixel_shuffle(c, 2)
        t = F.mish(gate) * torch.cat((ident, c), dim=1)
        t = rep(f'{p}.fc2', t) if f'{p}.fc2.alpha' in sd else _conv(sd, f'{p}.fc2', t, 0)
        out = F.mish(t) + short
    out = F.pixel_shuffle(rep('to_img.0', out), inner)
    return out[:, :, : h * scale, : w * scale] + F.interpolate(x, scale_factor=scale)


def gaterv3_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """GateRV3.forward in eval mode (/root/reference/resselt/archs/gaterv3/arch.py:783-802; MetaGated :658-667, GatedCNNBlock :623-629,
    InceptionDWConv2d :550-557, RMSNorm :518-524, SPAB :497-508, Conv3XC.update_params :432-463, Block :682-692, UniUpsampleV3 :241-373:
    the 'pixelshuffle', 'pixelshuffledirect', 'nearest+conv', 'transpose+conv', 'pa_up' and 'dysample' heads and the scale-1 conv; Attention :560-591)."""
    x = x.to(dtype)
    g = lambda k: sd[k].to(dtype)
    inp = x
    B, C, H, W = x.shape
    L = _seq_len(sd, 'gater_encode')
    pad = 2 ** L
    x = F.pad(x, (0, (pad - W % pad) % pad, 0, (pad - H % pad) % pad), 'reflect')

    def conv(name, t, padding=0, groups=1):
        return F.conv2d(t, g(f'{name}.weight'), g(f'{name}.bias') if f'{name}.bias' in sd else None, padding=padding, groups=groups)

    def c3xc(prefix, t):  # the merged eval-mode kernel: 1x1 -> 3x3 -> 1x1 plus the 1x1 skip (biases optional)
        w1, w2, w3 = g(f'{prefix}.conv.0.weight'), g(f'{prefix}.conv.1.weight'), g(f'{prefix}.conv.2.weight')
        k = F.conv2d(w1.flip(2, 3).permute(1, 0, 2, 3), w2, padding=2).flip(2, 3).permute(1, 0, 2, 3)
        k = F.conv2d(k.flip(2, 3).permute(1, 0, 2, 3), w3).flip(2, 3).permute(1, 0, 2, 3) + F.pad(g(f'{prefix}.sk.weight'), [1, 1, 1, 1])
        bias = None
        if f'{prefix}.conv.0.bias' in sd:
            b = (w2 * g(f'{prefix}.conv.0.bias').reshape(1, -1, 1, 1)).sum((1, 2, 3)) + g(f'{prefix}.conv.1.bias')
            bias = (w3 * b.reshape(1, -1, 1, 1)).sum((1, 2, 3)) + g(f'{prefix}.conv.2.bias') + g(f'{prefix}.sk.bias')
        return F.conv2d(t, k, bias, padding=1)

    def spab(prefix, t):
        o1 = F.silu(c3xc(f'{prefix}.c1_r', t))
        o3 = c3xc(f'{prefix}.c3_r', F.silu(c3xc(f'{prefix}.c2_r', o1)))
        return (o3 + t) * (torch.sigmoid(o3) - 0.5), o1

    def rms(p, t):
        r = t.norm(2, dim=1, keepdim=True) * t.shape[1] ** -0.5
        return g(f'{p}.scale')[:, None, None] * (t / (r + 1e-6)) + g(f'{p}.offset')[:, None, None]

    def gated_cnn(p, t):
        d = t.shape[1]
        hidden, gc = int(1.5 * d), int(d * 0.125)
        gate, ident, c = torch.split(conv(f'{p}.fc1', rms(f'{p}.norm', t)), [hidden, hidden - d, d], dim=1)
        if f'{p}.token_mix.qkv.weight' in sd:  # Attention (arch.py:572-591), 16 heads
            tm = f'{p}.token_mix'
            b_, c_, h_, w_ = c.shape
            heads = sd[f'{tm}.temperature'].shape[0]
            q, k, v = torch.chunk(conv(f'{tm}.qkv_dwconv', conv(f'{tm}.qkv', c), 1, 3 * c_), 3, dim=1)
            q, k, v = (u.reshape(b_, heads, c_ // heads, h_ * w_) for u in (q, k, v))
            attn = (F.normalize(q, dim=3) @ F.normalize(k, dim=3).transpose(2, 3)) * g(f'{tm}.temperature')
            c = conv(f'{tm}.project_out', (attn.softmax(dim=3) @ v).reshape(b_, c_, h_, w_))
        else:
            c_id, c_hw, c_w, c_h = torch.split(c, (d - 3 * gc, gc, gc, gc), dim=1)
            c = torch.cat((c_id, conv(f'{p}.token_mix.dwconv_hw', c_hw, 1, gc), conv(f'{p}.token_mix.dwconv_w', c_w, (0, 5), gc),
                           conv(f'{p}.token_mix.dwconv_h', c_h, (5, 0), gc)), dim=1)
        return F.mish(conv(f'{p}.fc2', F.mish(gate) * torch.cat((ident, c), dim=1)))

    def meta_gated(p, t):
        d = t.shape[1]
        u = conv(f'{p}.local.2', conv(f'{p}.local.1', rms(f'{p}.local.0', t)), 1, d)
        u1, u2 = u.chunk(2, dim=1)
        u = u1 * u2
        u = u * conv(f'{p}.sca.1', u.mean(dim=(2, 3), keepdim=True))
        u = u * g(f'{p}.gamma0') + t
        return gated_cnn(f'{p}.glob', u) * g(f'{p}.gamma1') + u

    x = conv('in_to_dim', x, 1)
    sisr, _ = spab('span_block0', x)
    sisr_short = sisr
    for k in range(_seq_len(sd, 'span_n_b')):
        sisr, _ = spab(f'span_n_b.{k}', sisr)
    sisr, sisr_out = spab('span_end', sisr)
    sisr = conv('sisr_cat_conv', torch.cat([x, c3xc('sisr_end_conv', sisr), sisr_short, sisr_out], dim=1))
    shorts = []
    for i in range(L):
        for j in range(_seq_len(sd, f'gater_encode.{i}.gated')):
            x = meta_gated(f'gater_encode.{i}.gated.{j}', x)
        shorts.append(x)
        x = F.pixel_unshuffle(conv(f'gater_encode.{i}.scale.0', x, 1), 2)
    for k in range(_seq_len(sd, 'latent')):
        x = gated_cnn(f'latent.{k}', x)
    for i in range(_seq_len(sd, 'decode')):
        x = torch.cat([F.pixel_shuffle(conv(f'decode.{i}.scale.0', x, 1), 2), shorts[L - 1 - i]], dim=1)
        x = conv(f'decode.{i}.shor', x)
        for j in range(_seq_len(sd, f'decode.{i}.gated')):
            x = meta_gated(f'decode.{i}.gated.{j}', x)
    x = x + sisr
    if 'dim_to_in.MetaUpsample' not in sd:
        scale, x = 1, conv('dim_to_in', x, 1)
    else:
        _, index, scale, _, _, _, groups = [int(v) for v in sd['dim_to_in.MetaUpsample']]
        mode = ('conv', 'pixelshuffledirect', 'pixelshuffle', 'nearest+conv', 'dysample', 'transpose+conv', 'lda', 'pa_up')[index]
        if mode == 'pixelshuffledirect':
            x = F.pixel_shuffle(conv('dim_to_in.0', x, 1), scale)
        elif mode == 'pixelshuffle':
            x = F.leaky_relu(conv('dim_to_in.0', x, 1), 0.01)
            i = 2
            for r in ([3] if scale == 3 else [2] * int(math.log2(scale))):
                x = F.pixel_shuffle(conv(f'dim_to_in.{i}', x, 1), r)
                i += 2
            x = conv(f'dim_to_in.{i}', x, 1)
        elif mode == 'nearest+conv':
            n, f = (1, 3) if scale == 3 else (int(math.log2(scale)), 2)
            for k in range(n):
                x = F.leaky_relu(F.interpolate(conv(f'dim_to_in.{3 * k}', x, 1), scale_factor=f), 0.2)
            x = conv(f'dim_to_in.{3 * n + 2}', F.leaky_relu(conv(f'dim_to_in.{3 * n}', x, 1), 0.2), 1)
        elif mode == 'transpose+conv':
            tconv = lambda name, t, s, pd: F.conv_transpose2d(t, g(f'{name}.weight'), g(f'{name}.bias'), stride=s, padding=pd)
            if scale == 4:
                x = conv('dim_to_in.3', tconv('dim_to_in.2', F.gelu(tconv('dim_to_in.0', x, 2, 1)), 2, 1), 1)
            else:
                x = conv('dim_to_in.1', tconv('dim_to_in.0', x, scale, 1 if scale == 2 else 0), 1)
        elif mode == 'pa_up':
            n, f = (1, 3) if scale == 3 else (int(math.log2(scale)), 2)
            for k in range(n):
                x = conv(f'dim_to_in.{6 * k + 1}', F.interpolate(x, scale_factor=f), 1)
                x = F.leaky_relu(x * torch.sigmoid(conv(f'dim_to_in.{6 * k + 2}.conv.0', x)), 0.2)
                x = F.leaky_relu(conv(f'dim_to_in.{6 * k + 4}', x, 1), 0.2)
            x = conv(f'dim_to_in.{6 * n}', x, 1)
        elif mode == 'dysample':
            i = 0
            if 'dim_to_in.0.weight' in sd:
                x, i = F.leaky_relu(conv('dim_to_in.0', x, 1), 0.01), 2
            x = _dysample(sd, f'dim_to_in.{i}', x, groups)
        else:
            raise NotImplementedError(mode)
    up = F.interpolate(inp, scale_factor=scale) if scale != 1 else inp
    return x[:, :, : H * scale, : W * scale] + g('gamma') * up


def _dysample(sd: SD, p: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """DySample.forward (/root/reference/resselt/utilities/dysample.py:46-83), restated with the same ATen calls: offsets
    ``offset(x) * sigmoid(scope(x)) * 0.5 + init_pos``, a coordinate grid in normalised [-1, 1] units, pixel_shuffle of the
    coordinates, ``grid_sample(bilinear, align_corners=False, padding_mode='border')`` per channel group, 1x1 ``end_conv``.
    (The reference builds its normaliser with ``pin_memory=True`` (:62), which needs a CUDA driver; the value is the same.)"""
    B, C, H, W = x.shape
    s = math.isqrt(sd[f'{p}.offset.weight'].shape[0] // (2 * groups))
    offset = _conv(sd, f'{p}.offset', x, 0) * torch.sigmoid(F.conv2d(x, sd[f'{p}.scope.weight'].to(x.dtype))) * 0.5 + sd[f'{p}.init_pos'].to(x.dtype)
    offset = offset.view(B, 2, -1, H, W)
    coords_h = torch.arange(H) + 0.5
    coords_w = torch.arange(W) + 0.5
    coords = torch.stack(torch.meshgrid([coords_w, coords_h], indexing='ij')).transpose(1, 2).unsqueeze(1).unsqueeze(0).type(x.dtype)
    normalizer = torch.tensor([W, H], dtype=x.dtype).view(1, 2, 1, 1, 1)
    coords = 2 * (coords + offset) / normalizer - 1
    coords = F.pixel_shuffle(coords.reshape(B, -1, H, W), s).view(B, 2, -1, s * H, s * W).permute(0, 2, 3, 4, 1).contiguous().flatten(0, 1)
    out = F.grid_sample(x.reshape(B * groups, -1, H, W), coords, mode='bilinear', align_corners=False, padding_mode='border').view(B, -1, s * H, s * W)
    if f'{p}.end_conv.weight' in sd:
        out = _conv(sd, f'{p}.end_conv', out, 0)
    return out


def spanplus_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """SpanPlus.forward with the 'ps' or 'dys' upsampler, eval mode
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
    if 'upsampler.0.weight' not in sd:  # 'dys' head: DySample(feature_channels, out_ch, upscale), groups = 4 (spanplus/arch.py:186)
        return _dysample(sd, 'upsampler', cur, 4)
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
    t = t + torch.repeat_interleave(x, r2, dim=1)
    if 'to_img.init_pos' in sd:  # DySample head: groups = out_ch for odd factors, else 4 (rplksr.py:133-140)
        r = math.isqrt(r2)
        return _dysample(sd, 'to_img', t, x.shape[1] if r % 2 != 0 else 4)
    return F.pixel_shuffle(t, math.isqrt(r2))


def plksr_forward(sd: SD, x: torch.Tensor, dtype=torch.float32, sparse_dilations=(1, 2, 3, 4), with_idt: bool = False) -> torch.Tensor:
    """plksr.forward (/root/reference/resselt/archs/plksr/plksr.py:374-377) with PLKBlock.forward (:301-308), the CCM / ICCM /
    DCCM mixers (:18-52, exact GELU), the three partial large-kernel layers in their un-converted eval form (PLKConv2d :70-81,
    SparsePLKConv2d :153-164 with the loader's default dilations, RectSparsePLKConv2d :116-117) and EA (:239-248)."""
    x = x.to(dtype)
    n_feats = _seq_len(sd, 'feats')
    t = _conv(sd, 'feats.0', x, 1)
    for i in range(1, n_feats - 1):
        p = f'feats.{i}'
        skip = t
        k0, k2 = sd[f'{p}.channe_mixer.0.weight'].shape[2], sd[f'{p}.channe_mixer.2.weight'].shape[2]
        t = _conv(sd, f'{p}.channe_mixer.2', F.gelu(_conv(sd, f'{p}.channe_mixer.0', t, k0 // 2)), k2 // 2)
        cw = lambda name, inp, **kw: F.conv2d(inp, sd[f'{p}.lk.{name}.weight'].to(dtype), sd[f'{p}.lk.{name}.bias'].to(dtype), **kw)
        if f'{p}.lk.conv.weight' in sd:
            pdim, k = sd[f'{p}.lk.conv.weight'].shape[0], sd[f'{p}.lk.conv.weight'].shape[2]
            x1 = t[:, :pdim]
            lk = cw('conv', x1, padding=k // 2)
        elif f'{p}.lk.convs.0.weight' in sd:
            pdim = sd[f'{p}.lk.convs.0.weight'].shape[0]
            x1 = t[:, :pdim]
            lk = 0.0
            for j in range(_seq_len(sd, f'{p}.lk.convs')):
                k = sd[f'{p}.lk.convs.{j}.weight'].shape[2]
                d = sparse_dilations[j] if j < len(sparse_dilations) else 1
                lk = lk + cw(f'convs.{j}', x1, padding=(k // 2) * d, dilation=d)
        else:
            pdim = sd[f'{p}.lk.mn_conv.weight'].shape[0]
            x1 = t[:, :pdim]
            m, n = sd[f'{p}.lk.mn_conv.weight'].shape[2:]
            lk = cw('mn_conv', x1, padding=(m // 2, n // 2)) + cw('nm_conv', x1, padding=(n // 2, m // 2)) + cw('nn_conv', x1, padding=n // 2)
        if with_idt:
            lk = lk + x1
        t = torch.cat([lk, t[:, pdim:]], 1)
        if f'{p}.attn.f.0.weight' in sd:
            t = t * torch.sigmoid(_conv(sd, f'{p}.attn.f.0', t, 1))
        t = _conv(sd, f'{p}.refine', t, 0) + skip
    t = _conv(sd, f'feats.{n_feats - 1}', t, 1)
    r2 = t.shape[1] // x.shape[1]
    return F.pixel_shuffle(t + torch.repeat_interleave(x, r2, dim=1), int(math.isqrt(r2)))


# ---------------------------------------------------------------------------------------------- DAT
def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    b = sd.get(name + '.bias')
    return F.linear(x, sd[name + '.weight'].to(x.dtype), None if b is None else b.to(x.dtype))


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + '.weight'].to(x.dtype), sd[name + '.bias'].to(x.dtype), 1e-5)


def _bn(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    g = lambda k: sd[f'{name}.{k}'].to(x.dtype)
    return F.batch_norm(x, g('running_mean'), g('running_var'), g('weight'), g('bias'), False, 0.0, 1e-5)


def _dat_pos_table(sd: SD, p: str, dtype) -> torch.Tensor:
    """DynamicPosBias.forward with residual=False (/root/reference/resselt/archs/dat/arch.py:135-143): a 4-layer MLP
    over the (2Hs-1)(2Ws-1) relative offsets stored in rpe_biases -> [offsets][heads]."""
    t = _lin(sd, f'{p}.pos.pos_proj', sd[f'{p}.rpe_biases'].to(dtype))
    for blk in ('pos1', 'pos2', 'pos3'):
        t = _lin(sd, f'{p}.pos.{blk}.2', F.relu(_ln(sd, f'{p}.pos.{blk}.0', t)))
    return t


def _dat_shift_mask(Hp, Wp, Hs, Ws, sh, sw, dtype):
    """calculate_mask (/root/reference/resselt/archs/dat/arch.py:363-428) for one branch: region labels of the rolled
    image, -100 between tokens of different regions inside a window."""
    img = torch.zeros(Hp, Wp, dtype=dtype)
    cnt = 0
    for hs in (slice(0, -Hs), slice(-Hs, -sh), slice(-sh, None)):
        for ws in (slice(0, -Ws), slice(-Ws, -sw), slice(-sw, None)):
            img[hs, ws] = cnt
            cnt += 1
    win = img.view(Hp // Hs, Hs, Wp // Ws, Ws).permute(0, 2, 1, 3).reshape(-1, Hs * Ws)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _dat_window_attention(sd: SD, p: str, q, k, v, Hp, Wp, Hs, Ws, heads, scale, mask):
    """Spatial_Attention.forward (/root/reference/resselt/archs/dat/arch.py:224-267); q,k,v: [B, Hp, Wp, C]."""
    B, _, _, C = q.shape

    def win(t):  # im2win (:217-222): -> [B*nW, heads, N, C/heads]
        t = t.view(B, Hp // Hs, Hs, Wp // Ws, Ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, Hs * Ws, heads, C // heads)
        return t.permute(0, 2, 1, 3)

    qw, kw, vw = win(q) * scale, win(k), win(v)
    attn = qw @ kw.transpose(-2, -1)
    table = _dat_pos_table(sd, p, q.dtype)
    idx = sd[f'{p}.relative_position_index'].long().view(-1)
    attn = attn + table[idx].view(Hs * Ws, Hs * Ws, -1).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B, nW, heads, Hs * Ws, Hs * Ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, Hs * Ws, Hs * Ws)
    out = (attn.softmax(-1) @ vw).transpose(1, 2).reshape(-1, Hs * Ws, C)
    return out.view(B, Hp // Hs, Wp // Ws, Hs, Ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)  # windows2img (:28-37)


def _dat_aim_convs(sd: SD, p: str, v_img):
    conv_x = F.gelu(_bn(sd, f'{p}.dwconv.1', F.conv2d(v_img, sd[f'{p}.dwconv.0.weight'].to(v_img.dtype), sd[f'{p}.dwconv.0.bias'].to(v_img.dtype),
                                                       padding=1, groups=v_img.shape[1])))
    def channel_interaction(t):
        t = F.adaptive_avg_pool2d(t, 1)
        t = F.gelu(_bn(sd, f'{p}.channel_interaction.2', _conv(sd, f'{p}.channel_interaction.1', t, 0)))
        return _conv(sd, f'{p}.channel_interaction.4', t, 0)
    def spatial_interaction(t):
        t = F.gelu(_bn(sd, f'{p}.spatial_interaction.1', _conv(sd, f'{p}.spatial_interaction.0', t, 0)))
        return _conv(sd, f'{p}.spatial_interaction.3', t, 0)
    return conv_x, channel_interaction, spatial_interaction


def _dat_spatial_block(sd: SD, p: str, x, H, W, split, shifted, heads):
    """Adaptive_Spatial_Attention.forward (/root/reference/resselt/archs/dat/arch.py:430-513)."""
    B, L, C = x.shape
    qkv = _lin(sd, f'{p}.qkv', x).reshape(B, L, 3, C).permute(2, 0, 1, 3)
    v_img = qkv[2].transpose(-2, -1).reshape(B, C, H, W)
    m = max(split)
    Hp, Wp = H + (m - H % m) % m, W + (m - W % m) % m
    qkv = F.pad(qkv.reshape(3 * B, H, W, C).permute(0, 3, 1, 2), (0, Wp - W, 0, Hp - H)).permute(0, 2, 3, 1).reshape(3, B, Hp, Wp, C)
    scale = (C // 2 // (heads // 2)) ** -0.5
    outs = []
    for br in (0, 1):
        Hs, Ws = (split[0], split[1]) if br == 0 else (split[1], split[0])
        sh, sw = Hs // 2, Ws // 2
        t = qkv[..., br * (C // 2):(br + 1) * (C // 2)]
        mask = None
        if shifted:
            t = torch.roll(t, shifts=(-sh, -sw), dims=(2, 3))
            mask = _dat_shift_mask(Hp, Wp, Hs, Ws, sh, sw, x.dtype)
        o = _dat_window_attention(sd, f'{p}.attns.{br}', t[0], t[1], t[2], Hp, Wp, Hs, Ws, heads // 2, scale, mask)
        if shifted:
            o = torch.roll(o, shifts=(sh, sw), dims=(1, 2))
        outs.append(o[:, :H, :W].reshape(B, L, C // 2))
    att = torch.cat(outs, 2)
    conv_x, channel_interaction, spatial_interaction = _dat_aim_convs(sd, p, v_img)
    cmap = channel_interaction(conv_x).permute(0, 2, 3, 1).reshape(B, 1, C)
    smap = spatial_interaction(att.transpose(-2, -1).reshape(B, C, H, W))
    att = att * torch.sigmoid(cmap)
    conv_x = (torch.sigmoid(smap) * conv_x).permute(0, 2, 3, 1).reshape(B, L, C)
    return _lin(sd, f'{p}.proj', att + conv_x)


def _dat_channel_block(sd: SD, p: str, x, H, W, heads):
    """Adaptive_Channel_Attention.forward (/root/reference/resselt/archs/dat/arch.py:565-612)."""
    B, N, C = x.shape
    qkv = _lin(sd, f'{p}.qkv', x).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = (t.transpose(-2, -1) for t in (qkv[0], qkv[1], qkv[2]))  # [B, heads, d, N]
    v_img = v.reshape(B, C, N).view(B, C, H, W)
    attn = (F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)) * sd[f'{p}.temperature'].to(x.dtype)
    att = (attn.softmax(-1) @ v).permute(0, 3, 1, 2).reshape(B, N, C)
    conv_x, channel_interaction, spatial_interaction = _dat_aim_convs(sd, p, v_img)
    cmap = channel_interaction(att.transpose(-2, -1).reshape(B, C, H, W))
    smap = spatial_interaction(conv_x).permute(0, 2, 3, 1).reshape(B, N, 1)
    att = att * torch.sigmoid(smap)
    conv_x = (conv_x * torch.sigmoid(cmap)).permute(0, 2, 3, 1).reshape(B, N, C)
    return _lin(sd, f'{p}.proj', att + conv_x)


def _dat_sgfn(sd: SD, p: str, x, H, W):
    """SGFN / SpatialGate (/root/reference/resselt/archs/dat/arch.py:40-101)."""
    B, N, _ = x.shape
    t = F.gelu(_lin(sd, f'{p}.fc1', x))
    x1, x2 = t.chunk(2, dim=-1)
    c = x2.shape[-1]
    x2 = _ln(sd, f'{p}.sg.norm', x2).transpose(1, 2).reshape(B, c, H, W)
    x2 = F.conv2d(x2, sd[f'{p}.sg.conv.weight'].to(x.dtype), sd[f'{p}.sg.conv.bias'].to(x.dtype), padding=1, groups=c)
    return _lin(sd, f'{p}.fc2', x1 * x2.flatten(2).transpose(-1, -2))


def dat_forward(sd: SD, x: torch.Tensor, dtype=torch.float32, img_range: float = 1.0) -> torch.Tensor:
    """DAT.forward, both heads ('pixelshuffle', 'pixelshuffledirect') and both residual connections ('1conv', '3conv') (/root/reference/resselt/archs/dat/arch.py:970-990,
    forward_features :959-968, ResidualGroup.forward :763-780, DATB.forward :674-683).  Which blocks shift follows
    :335 / :456: (rg even and b in {2, 6, ..}) or (rg odd and b % 4 == 0)."""
    x = x.to(dtype)
    in_ch = x.shape[1]
    mean = torch.tensor((0.4488, 0.4371, 0.4040), dtype=dtype).view(1, 3, 1, 1) if in_ch == 3 else torch.zeros(1, 1, 1, 1, dtype=dtype)
    x = (x - mean) * img_range
    B, _, H, W = x.shape
    split = [int(v) + 1 for v in sd['layers.0.blocks.0.attn.attns.0.rpe_biases'][-1]]
    feat = _conv(sd, 'conv_first', x, 1)
    t = _ln(sd, 'before_RG.1', feat.flatten(2).transpose(1, 2))
    for rg in range(_seq_len(sd, 'layers')):
        res = t
        for b in range(_seq_len(sd, f'layers.{rg}.blocks')):
            p = f'layers.{rg}.blocks.{b}'
            n1 = _ln(sd, f'{p}.norm1', t)
            if b % 2 == 0:
                heads = 2 * sd[f'{p}.attn.attns.0.pos.pos3.2.weight'].shape[0]
                shifted = (rg % 2 == 0 and b > 0 and (b - 2) % 4 == 0) or (rg % 2 != 0 and b % 4 == 0)
                t = t + _dat_spatial_block(sd, f'{p}.attn', n1, H, W, split, shifted, heads)
            else:
                t = t + _dat_channel_block(sd, f'{p}.attn', n1, H, W, sd[f'{p}.attn.temperature'].shape[0])
            t = t + _dat_sgfn(sd, f'{p}.ffn', _ln(sd, f'{p}.norm2', t), H, W)
        img = t.transpose(1, 2).reshape(B, -1, H, W)
        t = res + _swin_resi_conv(sd, f'layers.{rg}.conv', img).flatten(2).transpose(1, 2)  # '1conv' or '3conv' (:750-759)
    t = _ln(sd, 'norm', t).transpose(1, 2).reshape(B, -1, H, W)
    t = _swin_resi_conv(sd, 'conv_after_body', t) + feat
    if 'conv_last.weight' not in sd:  # 'pixelshuffledirect' head: UpsampleOneStep (:804-825, forward :981-984)
        r = math.isqrt(sd['upsample.0.weight'].shape[0] // in_ch)
        return F.pixel_shuffle(_conv(sd, 'upsample.0', t, 1), r) / img_range + mean
    t = F.leaky_relu(_conv(sd, 'conv_before_upsample.0', t, 1), 0.01)
    for i in range(0, _seq_len(sd, 'upsample'), 2):  # Upsample (:783-801): conv 64 -> 64 r^2 + PixelShuffle(r), r = 2 (n times) or 3
        wt = sd[f'upsample.{i}.weight']
        t = F.pixel_shuffle(_conv(sd, f'upsample.{i}', t, 1), math.isqrt(wt.shape[0] // wt.shape[1]))
    return _conv(sd, 'conv_last', t, 1) / img_range + mean


# ---------------------------------------------------------------------------------------------- SwinIR
def _swin_mask(H, W, ws, shift, dtype):
    """SwinTransformerBlock.calculate_mask (/root/reference/resselt/archs/swinir/arch.py:268-293)."""
    img = torch.zeros(H, W, dtype=dtype)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(H // ws, ws, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _swin_block(sd: SD, p: str, x, H, W, ws, shift, heads):
    """SwinTransformerBlock.forward (/root/reference/resselt/archs/swinir/arch.py:295-335) with WindowAttention.forward
    (:133-170) and Mlp.forward (:34-40) inlined; x: [B, H*W, C]."""
    B, L, C = x.shape
    t = _ln(sd, f'{p}.norm1', x).view(B, H, W, C)
    if shift > 0:
        t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
    N = ws * ws
    win = t.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, N, C)  # window_partition (:43-56)
    qkv = _lin(sd, f'{p}.attn.qkv', win).reshape(-1, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = sd[f'{p}.attn.relative_position_index'].long().view(-1)
    bias = sd[f'{p}.attn.relative_position_bias_table'].to(x.dtype)[idx].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift > 0:
        mask = _swin_mask(H, W, ws, shift, x.dtype)
        nW = mask.shape[0]
        attn = (attn.view(B, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, N, C)
    o = _lin(sd, f'{p}.attn.proj', o)
    o = o.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)  # window_reverse (:59-72)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.reshape(B, L, C)
    return x + _lin(sd, f'{p}.mlp.fc2', F.gelu(_lin(sd, f'{p}.mlp.fc1', _ln(sd, f'{p}.norm2', x))))


def _swin_resi_conv(sd: SD, name: str, t):
    """'1conv' (a single 3x3) or '3conv' (3x3 -> lrelu(0.2) -> 1x1 -> lrelu(0.2) -> 3x3), arch.py:564-574 / :890-901."""
    if f'{name}.weight' in sd:
        return _conv(sd, name, t, 1)
    t = F.leaky_relu(_conv(sd, f'{name}.0', t, 1), 0.2)
    t = F.leaky_relu(_conv(sd, f'{name}.2', t, 0), 0.2)
    return _conv(sd, f'{name}.4', t, 1)


def swinir_forward(sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """SwinIR.forward (/root/reference/resselt/archs/swinir/arch.py:962-1011; forward_features :947-960, RSTB.forward
    :592-593), all four reconstruction heads; hyper-parameters re-derived like the loader does
    (/root/reference/resselt/archs/swinir/__init__.py:35-96).  ape=False, patch_norm=True, start_unshuffle=1."""
    x = x.to(dtype)
    if 'conv_before_upsample.0.weight' in sd:
        upsampler = 'nearest+conv' if 'conv_up1.weight' in sd else 'pixelshuffle'
    else:
        upsampler = 'pixelshuffledirect' if 'upsample.0.weight' in sd else ''
    ws = int(math.isqrt(sd['layers.0.residual_group.blocks.0.attn.relative_position_index'].shape[0]))
    img_range = 255.0 if ws == 7 else 1.0
    B, in_ch, H0, W0 = x.shape
    x = F.pad(x, (0, (ws - W0 % ws) % ws, 0, (ws - H0 % ws) % ws), 'reflect') if (H0 % ws or W0 % ws) else x  # check_image_size
    H, W = x.shape[2:]
    mean = torch.tensor((0.4488, 0.4371, 0.4040), dtype=dtype).view(1, 3, 1, 1) if in_ch == 3 else torch.zeros(1, 1, 1, 1, dtype=dtype)
    x = (x - mean) * img_range
    feat = _conv(sd, 'conv_first', x, 1)
    t = _ln(sd, 'patch_embed.norm', feat.flatten(2).transpose(1, 2))
    for i in range(_seq_len(sd, 'layers')):
        res = t
        heads = sd[f'layers.{i}.residual_group.blocks.0.attn.relative_position_bias_table'].shape[1]
        for b in range(_seq_len(sd, f'layers.{i}.residual_group.blocks')):
            t = _swin_block(sd, f'layers.{i}.residual_group.blocks.{b}', t, H, W, ws, 0 if b % 2 == 0 else ws // 2, heads)
        t = _swin_resi_conv(sd, f'layers.{i}.conv', t.transpose(1, 2).reshape(B, -1, H, W)).flatten(2).transpose(1, 2) + res
    t = _ln(sd, 'norm', t).transpose(1, 2).reshape(B, -1, H, W)
    t = _swin_resi_conv(sd, 'conv_after_body', t) + feat
    up = 1
    if upsampler == 'pixelshuffle':
        t = F.leaky_relu(_conv(sd, 'conv_before_upsample.0', t, 1), 0.01)
        for i in range(0, _seq_len(sd, 'upsample'), 2):
            r = int(math.isqrt(sd[f'upsample.{i}.weight'].shape[0] // sd[f'upsample.{i}.weight'].shape[1]))
            t = F.pixel_shuffle(_conv(sd, f'upsample.{i}', t, 1), r)
            up *= r
        t = _conv(sd, 'conv_last', t, 1)
    elif upsampler == 'pixelshuffledirect':
        up = int(math.isqrt(sd['upsample.0.weight'].shape[0] // in_ch))
        t = F.pixel_shuffle(_conv(sd, 'upsample.0', t, 1), up)
    elif upsampler == 'nearest+conv':
        t = F.leaky_relu(_conv(sd, 'conv_before_upsample.0', t, 1), 0.01)
        for n in (1, 2, 3):
            if f'conv_up{n}.weight' in sd:
                t = F.leaky_relu(_conv(sd, f'conv_up{n}', F.interpolate(t, scale_factor=2, mode='nearest'), 1), 0.2)
                up *= 2
        t = _conv(sd, 'conv_last', F.leaky_relu(_conv(sd, 'conv_hr', t, 1), 0.2), 1)
    else:
        t = x + _conv(sd, 'conv_last', t, 1)
    t = t / img_range + mean
    return t[:, :, : H0 * up, : W0 * up]


_FORWARDS: Dict[str, Callable] = {
    'SwinIR': swinir_forward,
    'DAT': dat_forward,
    'RealPLKSR': realplksr_forward,
    'PLKSR': plksr_forward,
    'ESRGAN': esrgan_forward,
    'SPAN': span_forward,
    'SPANPlus': spanplus_forward,
    'Compact': compact_forward,
    'SpanPP': spanpp_forward,
    'RTMoSR': rtmosr_forward,
    'GateRV3': gaterv3_forward,
}


def forward_by_name(name: str, sd: SD, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    with torch.inference_mode():
        return _FORWARDS[name](sd, x, dtype)
