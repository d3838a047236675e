"""SPAN (Swift Parameter-free Attention Network) on the B200 engine.

Reference: /root/reference/resselt/archs/span/arch.py:183-250 (model), :157-180 (SPAB),
:59-154 (Conv3XC) and /root/reference/resselt/archs/span/__init__.py:10-55 (detection + loader).
"""
from __future__ import annotations

from typing import Mapping

import torch

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import pixelshuffle_scale
from ._common import conv3xc_specs, conv_specs, merge_conv3xc, merge_pointwise_into_conv


def emit_spab(pb: PlanBuilder, w, prefix: str, src, dst, t1, t2, act: int) -> None:
    """One SPAB block (span/arch.py:167-180): three merged 3x3 convs; the third fuses the gate
    ``(o3 + x) * (sigmoid(o3) - 0.5)``.  ``t1`` receives the *activated* first conv output, which is
    what the reference hands out as the block's second result (in-place activation)."""
    pb.conv(src, t1, *merge_conv3xc(w, f'{prefix}.c1_r'), act=act)
    pb.conv(t1, t2, *merge_conv3xc(w, f'{prefix}.c2_r'), act=act)
    pb.conv(t2, dst, *merge_conv3xc(w, f'{prefix}.c3_r'), combine=N.COMB_SPAB_GATE, res1=src)


class SPAN(EngineModule):
    def __init__(
        self,
        *,
        num_in_ch: int = 3,
        num_out_ch: int = 3,
        feature_channels: int = 48,
        upscale: int = 4,
        norm: bool = True,
        img_range: float = 255.0,
        rgb_mean=(0.4488, 0.4371, 0.4040),
        seed: int = 0,
    ):
        f = feature_channels
        specs = conv3xc_specs('conv_1', num_in_ch, f)
        for i in range(1, 7):
            for c in ('c1_r', 'c2_r', 'c3_r'):
                specs += conv3xc_specs(f'block_{i}.{c}', f, f)
        specs += conv_specs('conv_cat', 4 * f, f, 1)
        specs += conv3xc_specs('conv_2', f, f)
        specs += conv_specs('upsampler.0', f, num_out_ch * upscale * upscale, 3)
        if not norm:
            specs += [('no_norm', (1,), 'buffer_zeros')]
        super().__init__(specs, num_in_ch, num_out_ch, upscale, seed=seed)
        self.feature_channels = f
        self.norm = norm
        self.img_range = float(img_range)
        self.rgb_mean = tuple(float(m) for m in rgb_mean)
        if f % 8 != 0:
            raise ValueError('feature_channels must be a multiple of 8 for the planar-8 activation layout')

    @property
    def receptive_radius(self) -> int:
        # conv_1 (1) + 6 SPABs x 3 convs + conv_2 (1) + upsampler conv (1); conv_cat is 1x1
        return 1 + 6 * 3 + 1 + 1

    def build_plan(self, pb: PlanBuilder, w) -> None:
        f = self.feature_channels
        cat = pb.buffer(4 * f)  # [conv_1 out | conv_2 out | block_1 out | act(block_6.c1_r)] == the reference's torch.cat
        feat, tail, b1, o1_end = (cat.slice(i * f, f) for i in range(4))
        t1, t2, p0, p1 = (pb.buffer(f) for _ in range(4))
        norm_kw = {}
        if self.norm:
            if self.in_channels != 3:
                raise ValueError('SPAN input normalisation needs a 3-channel input (span/arch.py:232-234)')
            norm_kw = dict(in_mean=self.rgb_mean, in_scale=self.img_range)
        pb.conv(INPUT, feat, *merge_conv3xc(w, 'conv_1'), **norm_kw)
        emit_spab(pb, w, 'block_1', feat, b1, t1, t2, N.ACT_SILU)
        emit_spab(pb, w, 'block_2', b1, p0, t1, t2, N.ACT_SILU)
        emit_spab(pb, w, 'block_3', p0, p1, t1, t2, N.ACT_SILU)
        emit_spab(pb, w, 'block_4', p1, p0, t1, t2, N.ACT_SILU)
        emit_spab(pb, w, 'block_5', p0, p1, t1, t2, N.ACT_SILU)
        emit_spab(pb, w, 'block_6', p1, p0, o1_end, t2, N.ACT_SILU)
        pb.conv(p0, tail, *merge_conv3xc(w, 'conv_2'))
        # conv_cat (1x1, 192 -> 48) and the upsampler conv (3x3, 48 -> 3 r^2) have nothing between them (span/arch.py:247-248):
        # merged on the host into one 3x3 conv over the 192-channel concat (exact, border pixels included: border_bias) — one pass
        # over the concat instead of a 1 GB 1x1 pass plus a 3x3 pass (148 + 51 us -> one launch at 1080p)
        wm, bm, border = merge_pointwise_into_conv(w['conv_cat.weight'], w['conv_cat.bias'], w['upsampler.0.weight'], w['upsampler.0.bias'])
        pb.conv(cat, OUTPUT, wm, bm, ps=self.upscale, border_bias=border)


class SPANArch(Architecture[SPAN]):
    def __init__(self):
        super().__init__(
            uid='SPAN',
            detect=KeyCondition.has_all(
                'conv_1.sk.weight',
                'block_1.c1_r.sk.weight',
                'block_1.c1_r.eval_conv.weight',
                'block_1.c3_r.eval_conv.weight',
                'conv_cat.weight',
                'conv_2.sk.weight',
                'conv_2.eval_conv.weight',
                'upsampler.0.weight',
            ),
        )

    def load(self, state_dict: Mapping[str, object]):
        sk = state_dict['conv_1.sk.weight']
        num_in_ch, feature_channels = sk.shape[1], sk.shape[0]
        num_out_ch = num_in_ch
        upscale = pixelshuffle_scale(state_dict['upsampler.0.weight'].shape[0], num_in_ch)
        norm = 'no_norm' not in state_dict
        if not norm:
            # like the reference loader (span/__init__.py:41-43) the marker is normalised to zeros(1)
            state_dict['no_norm'] = torch.zeros(1)
        model = SPAN(
            num_in_ch=num_in_ch,
            num_out_ch=num_out_ch,
            feature_channels=feature_channels,
            upscale=upscale,
            norm=norm,
            img_range=255.0,  # not recoverable from the checkpoint (span/__init__.py:27-29)
            rgb_mean=(0.4488, 0.4371, 0.4040),
        )
        return self._enhance_model(model=model, in_channels=num_in_ch, out_channels=num_out_ch, upscale=upscale, name='SPAN')
