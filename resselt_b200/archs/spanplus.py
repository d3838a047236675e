"""SPANPlus on the B200 engine.

Reference: /root/reference/resselt/archs/spanplus/arch.py:154-201 (model), :133-151 (SPABS),
:105-130 (SPAB, Mish) and /root/reference/resselt/archs/spanplus/__init__.py:8-38 (loader).
Only the PixelShuffle ('ps') upsampler is in scope; DySample checkpoints are refused explicitly
(the reference's own DySample forward needs an NVIDIA driver, utilities/dysample.py:62).
"""
from __future__ import annotations

from typing import List, Mapping

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import dysample_scale, get_seq_len, pixelshuffle_scale
from ._common import merge_pointwise_into_conv, conv3xc_specs, conv_specs, dysample_specs, emit_dysample, merge_conv3xc
from .span import emit_spab


class SpanPlus(EngineModule):
    def __init__(
        self,
        num_in_ch: int = 3,
        num_out_ch: int = 3,
        blocks=(4,),
        feature_channels: int = 48,
        upscale: int = 4,
        drop_rate: float = 0.0,
        upsampler: str = 'ps',
        seed: int = 0,
    ):
        if upsampler not in ('ps', 'dys'):
            # 'conv' (1x only) cannot even be loaded by the reference (spanplus/__init__.py:20-27 reads upsampler.end_conv)
            raise NotImplementedError(f"SPANPlus upsampler '{upsampler}' is not supported by the B200 engine (only 'ps' and 'dys')")
        blocks = [int(blocks)] if not isinstance(blocks, (list, tuple)) else [int(b) for b in blocks]
        f = feature_channels
        specs = conv3xc_specs('feats.0', num_in_ch, f)
        for g, n_blocks in enumerate(blocks, start=1):
            names = ['block_1'] + [f'block_n.{i}' for i in range(n_blocks)] + ['block_end']
            for b in names:
                for c in ('c1_r', 'c2_r', 'c3_r'):
                    specs += conv3xc_specs(f'feats.{g}.{b}.{c}', f, f)
            specs += conv3xc_specs(f'feats.{g}.conv_2', f, f)
            specs += conv_specs(f'feats.{g}.conv_cat', 4 * f, f, 1)
        if upsampler == 'ps':
            out_ch = num_in_ch  # 'ps' models emit as many channels as they take (spanplus/arch.py:172)
            specs += conv_specs('upsampler.0', f, out_ch * upscale * upscale, 3)
        else:
            out_ch = num_out_ch
            specs += dysample_specs('upsampler', f, out_ch, upscale)  # DySample(feature_channels, out_channels, upscale) (spanplus/arch.py:186)
        super().__init__(specs, num_in_ch, out_ch, upscale, seed=seed)
        self.blocks: List[int] = blocks
        self.feature_channels = f
        self.upsampler_kind = upsampler
        if f % 8 != 0:
            raise ValueError('feature_channels must be a multiple of 8 for the planar-8 activation layout')

    @property
    def receptive_radius(self) -> int:
        if self.upsampler_kind != 'ps':
            raise NotImplementedError('DySample samples at learned, data-dependent offsets: no exact halo exists')
        # stem (1) + per group: (n + 2) SPABs x 3 convs + conv_2 (1); + upsampler conv (1)
        return 1 + sum(3 * (n + 2) + 1 for n in self.blocks) + 1

    def build_plan(self, pb: PlanBuilder, w) -> None:
        f = self.feature_channels
        t1, t2, p0, p1 = (pb.buffer(f) for _ in range(4))
        cats = [pb.buffer(4 * f) for _ in self.blocks]
        # group input lives in slot 0 of that group's concat buffer
        pb.conv(INPUT, cats[0].slice(0, f), *merge_conv3xc(w, 'feats.0'))
        for g, n_blocks in enumerate(self.blocks, start=1):
            cat = cats[g - 1]
            x_in, tail, b1, o1_end = (cat.slice(i * f, f) for i in range(4))
            pre = f'feats.{g}'
            emit_spab(pb, w, f'{pre}.block_1', x_in, b1, t1, t2, N.ACT_MISH)
            cur, pong = b1, [p0, p1]
            for i in range(n_blocks):
                nxt = pong[i % 2]
                emit_spab(pb, w, f'{pre}.block_n.{i}', cur, nxt, t1, t2, N.ACT_MISH)
                cur = nxt
            end_out = pong[n_blocks % 2]
            emit_spab(pb, w, f'{pre}.block_end', cur, end_out, o1_end, t2, N.ACT_MISH)
            pb.conv(end_out, tail, *merge_conv3xc(w, f'{pre}.conv_2'))  # Dropout2d is the identity in eval
            last = g == len(self.blocks)
            if last and self.upsampler_kind == 'ps':
                # the last group's conv_cat feeds the upsampler conv directly: one merged 3x3 conv over the concat (see span.py)
                wm, bm, border = merge_pointwise_into_conv(w[f'{pre}.conv_cat.weight'], w[f'{pre}.conv_cat.bias'], w['upsampler.0.weight'],
                                                           w['upsampler.0.bias'])
                pb.conv(cat, OUTPUT, wm, bm, ps=self.upscale, border_bias=border)
                return
            dst = cats[g].slice(0, f) if not last else t1
            pb.conv(cat, dst, w[f'{pre}.conv_cat.weight'], w[f'{pre}.conv_cat.bias'])
        if self.upsampler_kind == 'ps':
            pb.conv(t1, OUTPUT, w['upsampler.0.weight'], w['upsampler.0.bias'], ps=self.upscale)
        else:
            emit_dysample(pb, w, 'upsampler', t1, self.out_channels, self.upscale)


class SpanPlusArch(Architecture[SpanPlus]):
    def __init__(self):
        super().__init__(uid='spanplus', detect=KeyCondition.has_all('feats.0.eval_conv.weight'))

    def load(self, state_dict: Mapping[str, object]):
        n_groups = get_seq_len(state_dict, 'feats') - 1
        blocks = [get_seq_len(state_dict, f'feats.{g + 1}.block_n') for g in range(n_groups)]
        head = state_dict['feats.0.eval_conv.weight']
        num_in_ch, feature_channels = head.shape[1], head.shape[0]
        if 'upsampler.0.weight' in state_dict:
            upsampler, num_out_ch = 'ps', num_in_ch
            upscale = pixelshuffle_scale(state_dict['upsampler.0.weight'].shape[0], num_out_ch)
        else:
            upsampler = 'dys'
            num_out_ch = state_dict['upsampler.end_conv.weight'].shape[0]
            upscale = dysample_scale(state_dict['upsampler.offset.weight'].shape[0])
        model = SpanPlus(
            num_in_ch=num_in_ch,
            num_out_ch=num_out_ch,
            blocks=blocks,
            feature_channels=feature_channels,
            upscale=upscale,
            upsampler=upsampler,
        )
        return self._enhance_model(model=model, in_channels=num_in_ch, out_channels=num_out_ch, upscale=upscale, name='SPANPlus')
