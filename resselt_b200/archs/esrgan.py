"""ESRGAN / RRDBNet (old-arch, ESRGAN+, and new-arch Real-ESRGAN / BSRGAN key names) on the B200 engine.

Reference: /root/reference/resselt/archs/esrgan/arch.py:12-138 (model), /root/reference/resselt/utilities/block.py
:277-344 (RRDB), :347-465 (ResidualDenseBlock_5C), :510-537 (upconv_block), :148-200 (conv_block) and the loader
/root/reference/resselt/archs/esrgan/__init__.py:14-194.

What is fused where (per RDB, utilities/block.py:454-465):
  * the four growing ``torch.cat``s disappear: x, x1..x4 are channel ranges of one 192-channel planar buffer;
  * ``x5 * 0.2 + x`` (and, for the third RDB of an RRDB, ``out * 0.2 + x_rrdb``) is the last conv's epilogue;
  * ``nn.Upsample(x2, nearest)`` + 3x3 conv becomes four 2x2 phase convolutions on the low-resolution grid
    (the 3x3 taps that land on the same source pixel are pre-summed), 16 instead of 36 taps per output pixel.

Superset of the reference: checkpoints with new-arch key names (``conv_first``/``body.N.rdbM.convK`` or
``RRDB_trunk``/``trunk_conv``) load here; the reference converts them to a local dict it then drops
(esrgan/__init__.py:159 vs registry.py:112-113) and fails in the strict load.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Mapping

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder, Ref
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len
from ._common import conv_specs

GROWTH = 32  # gc is fixed to 32 by the reference constructor (esrgan/arch.py:86)


class _Keys:
    """Checkpoint key names of the three RRDBNet dialects, addressed by role."""

    def __init__(self, style: str, num_blocks: int, n_up: int):
        self.style, self.nb, self.n_up = style, num_blocks, n_up

    def first(self):
        return 'model.0' if self.style == 'old' else 'conv_first'

    def rdb_conv(self, i: int, j: int, k: int):
        if self.style == 'old':
            return f'model.1.sub.{i}.RDB{j}.conv{k}.0'
        if self.style == 'new':
            return f'body.{i}.rdb{j}.conv{k}'
        return f'RRDB_trunk.{i}.RDB{j}.conv{k}'

    def rdb_conv1x1(self, i: int, j: int):
        return f'model.1.sub.{i}.RDB{j}.conv1x1'

    def trunk(self):
        return {'old': f'model.1.sub.{self.nb}', 'new': 'conv_body', 'bsrgan': 'trunk_conv'}[self.style]

    def up(self, u: int):  # u = 1..n_up
        return {'old': f'model.{3 * u}', 'new': f'conv_up{u}', 'bsrgan': f'upconv{u}'}[self.style]

    def hr(self):
        return {'old': f'model.{3 * self.n_up + 2}', 'new': 'conv_hr', 'bsrgan': 'HRconv'}[self.style]

    def last(self):
        return f'model.{3 * self.n_up + 4}' if self.style == 'old' else 'conv_last'


def upconv_phase_kernels(w: torch.Tensor):
    """nearest-x2 upsample followed by a 3x3 'same' conv == four 2x2 convs on the source grid.

    For output pixel (2y+a, 2x+b) the tap ky reads upsampled row 2y+a+ky-1, i.e. source row y + ((a+ky-1) >> 1):
    phase a=0 folds ky=1,2 onto row y (ky=0 -> y-1); phase a=1 folds ky=0,1 onto row y (ky=2 -> y+1).
    Returns [(phase index a*2+b, weight [cout][cin][2][2], (pad_top, pad_left))]."""
    rows = {0: ([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1), 1: ([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], 0)}
    out = []
    for a in (0, 1):
        (r0, r1), pad_t = rows[a]
        for b in (0, 1):
            def fold(r):  # r: [cout][cin][3] over kx
                return (torch.stack([r[:, :, 0], r[:, :, 1] + r[:, :, 2]], -1), 1) if b == 0 else (
                    torch.stack([r[:, :, 0] + r[:, :, 1], r[:, :, 2]], -1), 0)
            (c0, pad_l), (c1, _) = fold(r0), fold(r1)
            out.append((a * 2 + b, torch.stack([c0, c1], 2).contiguous(), (pad_t, pad_l)))
    return out


class RRDBNet(EngineModule):
    def __init__(
        self,
        in_nc: int = 3,
        out_nc: int = 3,
        num_filters: int = 64,
        num_blocks: int = 23,
        scale: int = 4,
        plus: bool = False,
        shuffle_factor: int | None = None,
        key_style: str = 'old',
        seed: int = 0,
    ):
        if scale & (scale - 1) or scale < 1:
            raise NotImplementedError('only power-of-two scales (nearest-x2 upconv stacks) are supported')
        if plus and key_style != 'old':
            raise ValueError('ESRGAN+ checkpoints only exist with old-arch keys')
        n_up = int(math.log2(scale))
        nf, gc = num_filters, GROWTH
        keys = _Keys(key_style, num_blocks, n_up)
        specs = conv_specs(keys.first(), in_nc, nf, 3)
        for i in range(num_blocks):
            for j in (1, 2, 3):
                if plus:
                    specs += [(keys.rdb_conv1x1(i, j) + '.weight', (gc, nf, 1, 1), 'conv_w')]
                for k in range(1, 6):
                    specs += conv_specs(keys.rdb_conv(i, j, k), nf + (k - 1) * gc, gc if k < 5 else nf, 3)
        specs += conv_specs(keys.trunk(), nf, nf, 3)
        for u in range(1, n_up + 1):
            specs += conv_specs(keys.up(u), nf, nf, 3)
        specs += conv_specs(keys.hr(), nf, nf, 3)
        specs += conv_specs(keys.last(), nf, out_nc, 3)
        self.shuffle_factor = shuffle_factor
        upscale = scale // shuffle_factor if shuffle_factor else scale
        visible_in = in_nc // (shuffle_factor**2) if shuffle_factor else in_nc
        super().__init__(specs, visible_in, out_nc, upscale, seed=seed, plan_io=(in_nc, out_nc, scale))
        self._keys = keys
        self.num_filters, self.num_blocks, self.n_up, self.plus = nf, num_blocks, n_up, plus
        if nf % 8 != 0:
            raise ValueError('num_filters must be a multiple of 8 for the planar-8 activation layout')

    @property
    def receptive_radius(self) -> int:
        # plan-grid pixels: first conv (1) + 15 convs per RRDB + trunk conv (1) + the HR-side convs (2 LR pixels cover them);
        # behind a pixel-unshuffle front end (Real-ESRGAN x2 / x1) one plan-grid pixel is shuffle_factor caller pixels
        return (15 * self.num_blocks + 4) * (self.shuffle_factor or 1)

    @property
    def tile_multiple(self) -> int:
        # tiles must start on multiples of shuffle_factor: otherwise pixel_unshuffle phases (and the reflect pad of odd sizes,
        # esrgan/arch.py:130-137) differ from the untiled forward
        return self.shuffle_factor or 1

    # ------------------------------------------------------------------ plan
    def _conv5(self, pb: PlanBuilder, w, name: str, src: Ref, dst: Ref, tmp: Ref, alpha: float, r_in: Ref, r_outer: Ref | None):
        """Last conv of an RDB with its residual tail:  dst = alpha * conv(src) + beta_in * r_in [+ r_outer].
        On the tensor-core path the 192->64 kernel (221 KB in bf16) does not fit shared memory next to the halo
        tile, so the contraction is split over the input channels: the first half leaves a partial result in ``tmp``."""
        wt, bias = w[name + '.weight'], w[name + '.bias']
        cin = wt.shape[1]
        beta_in = 0.2 if r_outer is not None else 1.0
        fits = pb.compute_dtype != torch.bfloat16 or wt.shape[0] * cin * 9 * 2 <= 150 * 1024
        if fits:
            pb.conv(src, dst, wt, bias, combine=N.COMB_AXPY, alpha=alpha, res1=r_in, beta1=beta_in, res2=r_outer, beta2=1.0)
            return
        half = (cin // 2 + 15) // 16 * 16
        pb.conv(src.slice(0, half), tmp, wt[:, :half], None, combine=N.COMB_AXPY, alpha=alpha, res1=r_in, beta1=beta_in)
        pb.conv(src.slice(half, cin - half), dst, wt[:, half:], bias, combine=N.COMB_AXPY, alpha=alpha, res1=tmp, beta1=1.0,
                res2=r_outer, beta2=1.0)

    def _rdb(self, pb: PlanBuilder, w, i: int, j: int, dense: Ref, dst: Ref, tmp: Ref, plus_tmp, r_outer: Ref | None):
        nf, gc, keys = self.num_filters, GROWTH, self._keys
        lrelu = dict(act=N.ACT_LRELU, act_param=0.2)
        x = dense.slice(0, nf)
        grow = [dense.slice(nf + k * gc, gc) for k in range(4)]
        name = lambda k: keys.rdb_conv(i, j, k)
        pb.conv(x, grow[0], w[name(1) + '.weight'], w[name(1) + '.bias'], **lrelu)
        if self.plus:
            # ESRGAN+: x2 = lrelu(conv2) + conv1x1(x);  x4 = lrelu(conv4) + x2   (utilities/block.py:457-463)
            pb.conv(x, plus_tmp, w[keys.rdb_conv1x1(i, j) + '.weight'], None)
            pb.conv(dense.slice(0, nf + gc), grow[1], w[name(2) + '.weight'], w[name(2) + '.bias'], combine=N.COMB_AXPY, res1=plus_tmp, **lrelu)
        else:
            pb.conv(dense.slice(0, nf + gc), grow[1], w[name(2) + '.weight'], w[name(2) + '.bias'], **lrelu)
        pb.conv(dense.slice(0, nf + 2 * gc), grow[2], w[name(3) + '.weight'], w[name(3) + '.bias'], **lrelu)
        if self.plus:
            pb.conv(dense.slice(0, nf + 3 * gc), grow[3], w[name(4) + '.weight'], w[name(4) + '.bias'], combine=N.COMB_AXPY, res1=grow[1], **lrelu)
        else:
            pb.conv(dense.slice(0, nf + 3 * gc), grow[3], w[name(4) + '.weight'], w[name(4) + '.bias'], **lrelu)
        # x5 * 0.2 + x ; for the RRDB's last RDB also (...) * 0.2 + x_rrdb  ->  0.04 * x5 + 0.2 * x + x_rrdb
        alpha = 0.04 if r_outer is not None else 0.2
        self._conv5(pb, w, name(5), dense, dst, tmp, alpha, x, r_outer)

    def build_plan(self, pb: PlanBuilder, w) -> None:
        nf, gc, keys = self.num_filters, GROWTH, self._keys
        lrelu = dict(act=N.ACT_LRELU, act_param=0.2)
        fea = pb.buffer(nf)
        dense = [pb.buffer(nf + 4 * gc) for _ in range(3)]
        tmp = pb.buffer(nf)
        plus_tmp = pb.buffer(gc) if self.plus else None
        first = keys.first()
        pb.conv(INPUT, fea, w[first + '.weight'], w[first + '.bias'])
        pb.conv(INPUT, dense[0].slice(0, nf), w[first + '.weight'], w[first + '.bias'])  # RRDB 0 needs it inside a dense buffer
        a, b, c = dense
        for i in range(self.num_blocks):
            self._rdb(pb, w, i, 1, a, b.slice(0, nf), tmp, plus_tmp, None)
            self._rdb(pb, w, i, 2, b, c.slice(0, nf), tmp, plus_tmp, None)
            self._rdb(pb, w, i, 3, c, b.slice(0, nf), tmp, plus_tmp, a.slice(0, nf))
            a, b, c = b, c, a
        trunk = keys.trunk()
        cur = tmp  # ShortcutBlock: fea + trunk_conv(rrdb stack)
        pb.conv(a.slice(0, nf), cur, w[trunk + '.weight'], w[trunk + '.bias'], combine=N.COMB_AXPY, res1=fea)
        grid = 1
        for u in range(1, self.n_up + 1):
            nxt = pb.buffer(nf, scale=grid * 2)
            name = keys.up(u)
            for phase, wk, pad in upconv_phase_kernels(w[name + '.weight']):
                pb.conv(cur, nxt, wk, w[name + '.bias'], dst_ps=2, dst_phase=phase, pad=pad, **lrelu)
            cur, grid = nxt, grid * 2
        hr = pb.buffer(nf, scale=grid)
        pb.conv(cur, hr, w[keys.hr() + '.weight'], w[keys.hr() + '.bias'], **lrelu)
        pb.conv(hr, OUTPUT, w[keys.last() + '.weight'], w[keys.last() + '.bias'], ps=1)

    # ------------------------------------------------------------------ Real-ESRGAN x2/x1 front end
    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if not self.shuffle_factor:
            return super().forward_into(x, out)
        out.copy_(self.forward(x))
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.shuffle_factor:
            return super().forward(x)
        # host-side glue of esrgan/arch.py:130-137: reflect-pad to a multiple of the factor, pixel-unshuffle, crop
        f = self.shuffle_factor
        h, w = x.shape[-2:]
        x = F.pad(x, (0, (f - w % f) % f, 0, (f - h % f) % f), 'reflect')
        y = super().forward(F.pixel_unshuffle(x, f).contiguous())
        return y[:, :, : h * self.upscale, : w * self.upscale]


_NEW_RDB = re.compile(r'^body\.(\d+)\.rdb(\d)\.conv(\d+)\.(weight|bias)$')
_BSR_RDB = re.compile(r'^RRDB_trunk\.(\d+)\.RDB(\d)\.conv(\d+)\.(weight|bias)$')


class ESRGANArch(Architecture[RRDBNet]):
    def __init__(self) -> None:
        super().__init__(
            uid='ESRGAN',
            detect=KeyCondition.has_any(
                KeyCondition.has_all('model.0.weight', 'model.1.sub.0.RDB1.conv1.0.weight'),
                KeyCondition.has_all('conv_first.weight', 'body.0.rdb1.conv1.weight', 'conv_body.weight', 'conv_last.weight'),
                KeyCondition.has_all('conv_first.weight', 'RRDB_trunk.0.RDB1.conv1.weight', 'trunk_conv.weight', 'conv_last.weight'),
                KeyCondition.has_all('model.0.weight', 'model.1.sub.0.RDB1.conv1x1.weight'),
            ),
        )

    def load(self, state_dict: Mapping[str, object]):
        if 'model.0.weight' in state_dict:
            style = 'old'
            seq_len = get_seq_len(state_dict, 'model')
            in_nc = state_dict['model.0.weight'].shape[1]
            out_nc = state_dict[f'model.{seq_len - 1}.weight'].shape[0]
            scale = 2 ** ((seq_len - 5) // 3)  # [Upsample, Conv, LeakyReLU] triples between the fixed 5 layers
            num_blocks = get_seq_len(state_dict, 'model.1.sub') - 1
            num_filters = state_dict['model.0.weight'].shape[0]
        else:
            style = 'new' if 'conv_body.weight' in state_dict else 'bsrgan'
            pattern = _NEW_RDB if style == 'new' else _BSR_RDB
            num_blocks = 1 + max(int(m.group(1)) for m in map(pattern.match, state_dict) if m)
            in_nc = state_dict['conv_first.weight'].shape[1]
            num_filters = state_dict['conv_first.weight'].shape[0]
            out_nc = state_dict['conv_last.weight'].shape[0]
            up = 'conv_up' if style == 'new' else 'upconv'
            scale = 2 ** sum(1 for u in range(1, 6) if f'{up}{u}.weight' in state_dict)
        plus = any('.conv1x1.' in k for k in state_dict)
        shuffle_factor = None
        if in_nc in (out_nc * 4, out_nc * 16):  # Real-ESRGAN x2 / x1 models pixel-unshuffle their input
            shuffle_factor = int(math.sqrt(in_nc / out_nc))
        model = RRDBNet(in_nc=in_nc, out_nc=out_nc, num_filters=num_filters, num_blocks=num_blocks, scale=scale, plus=plus,
                        shuffle_factor=shuffle_factor, key_style=style)
        if shuffle_factor:
            in_nc //= shuffle_factor**2
            scale //= shuffle_factor
        return self._enhance_model(model=model, in_channels=in_nc, out_channels=out_nc, upscale=scale, name='ESRGAN')
