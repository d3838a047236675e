"""Compact (SRVGGNetCompact) on the B200 engine.

Reference: /root/reference/resselt/archs/compact/arch.py:5-65 and
/root/reference/resselt/archs/compact/__init__.py:8-38.
"""
from __future__ import annotations

from typing import Mapping

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len, pixelshuffle_scale
from ._common import conv_specs


class SRVGGNetCompact(EngineModule):
    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type='prelu', seed: int = 0):
        if act_type != 'prelu':
            raise NotImplementedError('only the PReLU variant is loadable through the registry (compact/arch.py:40)')
        specs = []
        chans = [num_in_ch] + [num_feat] * (num_conv + 1)
        for i in range(num_conv + 1):  # conv at body.{2i}, PReLU at body.{2i+1}
            specs += conv_specs(f'body.{2 * i}', chans[i], num_feat, 3)
            specs += [(f'body.{2 * i + 1}.weight', (num_feat,), 'prelu')]
        specs += conv_specs(f'body.{2 * num_conv + 2}', num_feat, num_out_ch * upscale * upscale, 3)
        super().__init__(specs, num_in_ch, num_out_ch, upscale, seed=seed)
        self.num_feat, self.num_conv = num_feat, num_conv
        if num_in_ch != num_out_ch:
            raise ValueError('the nearest-upsampled input residual needs num_in_ch == num_out_ch (compact/arch.py:63-64)')

    @property
    def receptive_radius(self) -> int:
        return self.num_conv + 2  # every body conv is a 3x3

    def build_plan(self, pb: PlanBuilder, w) -> None:
        ping = [pb.buffer(self.num_feat), pb.buffer(self.num_feat)]
        src = INPUT
        for i in range(self.num_conv + 1):
            dst = ping[i % 2]
            pb.conv(src, dst, w[f'body.{2 * i}.weight'], w[f'body.{2 * i}.bias'], act=N.ACT_PRELU, act_slopes=w[f'body.{2 * i + 1}.weight'])
            src = dst
        last = 2 * self.num_conv + 2
        # PixelShuffle + "out += nearest_upsample(x)" both happen in the last conv's store
        pb.conv(src, OUTPUT, w[f'body.{last}.weight'], w[f'body.{last}.bias'], ps=self.upscale, add_base=True)


class CompactArch(Architecture[SRVGGNetCompact]):
    def __init__(self):
        super().__init__(uid='Compact', detect=KeyCondition.has_all('body.0.weight', 'body.1.weight'))

    def load(self, state_dict: Mapping[str, object]):
        top = get_seq_len(state_dict, 'body') - 1
        first = state_dict['body.0.weight']
        in_nc, num_feat = first.shape[1], first.shape[0]
        num_conv = (top - 2) // 2
        scale = pixelshuffle_scale(state_dict[f'body.{top}.bias'].shape[0], in_nc)
        model = SRVGGNetCompact(num_in_ch=in_nc, num_out_ch=in_nc, num_feat=num_feat, num_conv=num_conv, upscale=scale)
        return self._enhance_model(model=model, in_channels=in_nc, out_channels=in_nc, upscale=scale, name='Compact')
