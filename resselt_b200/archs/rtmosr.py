"""RTMoSR on the B200 engine — SURVEY.md section 8f rank 2, the second of the SPAN descendants.

Reference: /root/reference/resselt/archs/rtmosr/arch.py:329-387 (model), :295-326 (GatedCNNBlock: RMSNorm -> RepConv fc1 ->
split into gate / identity / conv parts -> [ParPixelUnshuffle -> OmniShift -> CSELayer -> PixelShuffle] on the conv part ->
``mish(fc2(mish(g) * cat(i, c))) + shortcut``), :162-212 (RepConv, as in SpanPP), :215-281 (OmniShift: identity + depthwise 1x1 /
3x3 / 5x5 re-parameterised into one depthwise 5x5 by ``.eval()``), :284-292 (ParPixelUnshuffle), :7-21 (CSELayer), :25-38 (RMSNorm),
loader /root/reference/resselt/archs/rtmosr/__init__.py:9-104.

Lowering: every RepConv / OmniShift is merged once per plan on the host (fp64; the checkpoint's ``conv_3x3_rep.*``, ``eval_conv.*`` and
``conv5x5_reparam.*`` are dead values: ``.eval()`` overwrites them).  fc1 runs as three convs (identity part, conv part, gate part)
so that the gate conv's epilogue does ``mish(g) * cat(i, c)`` (COMB_MUL) once the conv branch has filled its half of the concat
buffer; fc2's epilogue does ``mish(.) + shortcut``; the last conv scatters PixelShuffle'd into the caller's tensor and adds the
nearest-upsampled input (``add_base``).  The half-resolution branch lives on a plan with base divisor 2: one op produces
PixelUnshuffle(2) and MaxPool2d(2) of the conv part, the pooled RepConv adds the unshuffled map in its epilogue, a depthwise 5x5 op,
then squeeze-excitation + PixelShuffle(2) in one op (rsb_op_kind 7-9).  Reflect padding to an even size, the optional
pixel-unshuffle front end and the final crop are host glue around the plan, like Real-ESRGAN's.
"""
from __future__ import annotations

import math
from typing import List, Mapping

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, ParamSpec, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len
from ._common import conv_specs
from .spanpp import merge_repconv, repconv_specs


def omnishift_specs(prefix: str, dim: int) -> List[ParamSpec]:
    specs: List[ParamSpec] = [(f'{prefix}.alpha{k}', (1, dim, 1, 1), 'affine_w') for k in (1, 2, 3, 4)]
    for name, k in (('conv1x1', 1), ('conv3x3', 3), ('conv5x5', 5)):
        specs += [(f'{prefix}.{name}.weight', (dim, 1, k, k), 'conv_w*2.0'), (f'{prefix}.{name}.bias', (dim,), f'bias:{k * k}')]
    specs += [(f'{prefix}.conv5x5_reparam.weight', (dim, 1, 5, 5), 'conv_w'), (f'{prefix}.conv5x5_reparam.bias', (dim,), 'bias:25')]
    return specs


def merge_omnishift(w, prefix: str):
    """OmniShift.reparam_5x5 (rtmosr/arch.py:249-269) in fp64: alpha1 * identity + alpha2 * dw1x1 + alpha3 * dw3x3 + alpha4 * dw5x5 as one
    depthwise 5x5 kernel [C][1][5][5] and bias [C]."""
    f64 = torch.float64
    g = lambda k: w[f'{prefix}.{k}'].to(f64)
    a1, a2, a3, a4 = (g(f'alpha{k}').reshape(-1, 1, 1, 1) for k in (1, 2, 3, 4))
    w1, w3, w5 = F.pad(g('conv1x1.weight'), (2, 2, 2, 2)), F.pad(g('conv3x3.weight'), (1, 1, 1, 1)), g('conv5x5.weight')
    ident = F.pad(torch.ones_like(g('conv1x1.weight')), (2, 2, 2, 2))
    weight = a1 * ident + a2 * w1 + a3 * w3 + a4 * w5
    bias = a2.flatten() * g('conv1x1.bias') + a3.flatten() * g('conv3x3.bias') + a4.flatten() * g('conv5x5.bias')
    return weight, bias


class RTMoSR(EngineModule):
    def __init__(self, *, scale: int = 2, dim: int = 32, ffn_expansion: float = 2, n_blocks: int = 2, unshuffle_mod: bool = False,
                 dccm: bool = True, se: bool = True, seed: int = 0):
        unshuffle, inner = 0, int(scale)
        if scale < 4 and unshuffle_mod:
            if scale == 3:
                raise ValueError('Unshuffle_mod does not support 3x')  # rtmosr/arch.py:345-346
            unshuffle, inner = 4 // int(scale), 4
        hidden = int(ffn_expansion * dim)
        if dim % 8 or hidden % 8 or hidden < dim:
            raise ValueError('dim and the hidden width must be multiples of 8 (planar-8 activation layout), hidden >= dim')
        feat = 'to_feat.1' if unshuffle else 'to_feat'
        specs = repconv_specs(feat, 3 * max(unshuffle, 1) ** 2, dim)
        for i in range(n_blocks):
            p = f'body.{i}'
            specs += [(f'{p}.norm.scale', (dim,), 'affine_w'), (f'{p}.norm.offset', (dim,), 'normal:0.1')]
            specs += repconv_specs(f'{p}.fc1', dim, 2 * hidden)
            specs += repconv_specs(f'{p}.conv.0.poll.1', dim, 4 * dim)
            specs += omnishift_specs(f'{p}.conv.1', 4 * dim)
            if se:
                specs += conv_specs(f'{p}.conv.2.squeezing.0', 4 * dim, 2 * dim, 1, gain=3.0) + conv_specs(f'{p}.conv.2.squeezing.2', 2 * dim, 4 * dim, 1, gain=3.0)
            specs += repconv_specs(f'{p}.fc2', hidden, dim) if dccm else conv_specs(f'{p}.fc2', hidden, dim, 1)
        specs += repconv_specs('to_img.0', dim, 3 * inner * inner)
        plan_in = 3 * max(unshuffle, 1) ** 2
        super().__init__(specs, 3, 3, int(scale), seed=seed, plan_io=(plan_in, 3, inner))
        self.scale = int(scale)
        self._plan_base_divisor = 2
        self.dim, self.hidden, self.n_blocks, self.dccm, self.se = dim, hidden, n_blocks, dccm, se
        self.unshuffle, self.inner_scale, self.ffn_expansion = unshuffle, inner, ffn_expansion
        self.pad = 2 * (unshuffle if unshuffle > 0 else 1)  # rtmosr/arch.py:349-350
        self._feat = feat

    # ------------------------------------------------------------------ host glue (arch.py:375-387)
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        xp = F.pad(x, (0, (self.pad - w % self.pad) % self.pad, 0, (self.pad - h % self.pad) % self.pad), 'reflect')
        if not self.unshuffle:
            # the plan's last conv adds the nearest-upsampled (padded) input itself; the crop keeps the part that belongs to x
            y = super().forward(xp.contiguous())
            return y if y.shape[-2:] == (h * self.scale, w * self.scale) else y[:, :, : h * self.scale, : w * self.scale].contiguous()
        y = super().forward(F.pixel_unshuffle(xp, self.unshuffle).contiguous())
        return y[:, :, : h * self.scale, : w * self.scale] + F.interpolate(x, scale_factor=self.scale)

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        if not self.unshuffle and h % self.pad == 0 and w % self.pad == 0:
            return super().forward_into(x, out)
        out.copy_(self.forward(x))
        return out

    # ------------------------------------------------------------------ plan
    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, hidden = self.dim, self.hidden
        x, x2, xn = pb.buffer(dim), pb.buffer(dim), pb.buffer(dim)           # full resolution (scale 2 on the half-grid plan)
        ic, gm = pb.buffer(hidden), pb.buffer(hidden)                         # cat(i, c) and mish(g) * cat(i, c)
        cpart = pb.buffer(dim)
        u5, v, o = pb.buffer(5 * dim, scale=1), pb.buffer(4 * dim, scale=1), pb.buffer(4 * dim, scale=1)   # half resolution
        pb.conv(INPUT, x, *merge_repconv(w, self._feat))
        for i in range(self.n_blocks):
            p = f'body.{i}'
            pb.rmsnorm(x, xn, w[f'{p}.norm.scale'], w[f'{p}.norm.offset'], eps=1e-6)
            w1, b1 = merge_repconv(w, f'{p}.fc1')
            # torch.split(fc1(x), [hidden, hidden - dim, dim], dim=1) (arch.py:311,322)
            g_rows, i_rows, c_rows = slice(0, hidden), slice(hidden, 2 * hidden - dim), slice(2 * hidden - dim, 2 * hidden)
            if hidden > dim:
                pb.conv(xn, ic.slice(0, hidden - dim), w1[i_rows], b1[i_rows])
            pb.conv(xn, cpart, w1[c_rows], b1[c_rows])
            # ParPixelUnshuffle: pu(c) + RepConv(MaxPool2d(2)(c))  (arch.py:284-292)
            pb.unshuffle_pool(cpart, u5)
            pb.conv(u5.slice(4 * dim, dim), v, *merge_repconv(w, f'{p}.conv.0.poll.1'), combine=N.COMB_AXPY, res1=u5.slice(0, 4 * dim))
            pb.dwconv(v, o, *merge_omnishift(w, f'{p}.conv.1'))
            se = None
            if self.se:
                se = (w[f'{p}.conv.2.squeezing.0.weight'], w[f'{p}.conv.2.squeezing.0.bias'], w[f'{p}.conv.2.squeezing.2.weight'],
                      w[f'{p}.conv.2.squeezing.2.bias'])
            pb.se_shuffle(o, ic.slice(hidden - dim, dim), se=se)
            pb.conv(xn, gm, w1[g_rows], b1[g_rows], act=N.ACT_MISH, combine=N.COMB_MUL, res1=ic)          # mish(g) * cat(i, c)
            w2, b2 = merge_repconv(w, f'{p}.fc2') if self.dccm else (w[f'{p}.fc2.weight'], w[f'{p}.fc2.bias'])
            pb.conv(gm, x2, w2, b2, act=N.ACT_MISH, combine=N.COMB_AXPY, res1=x)                          # mish(fc2(.)) + shortcut
            x, x2 = x2, x
        pb.conv(x, OUTPUT, *merge_repconv(w, 'to_img.0'), ps=self.inner_scale, add_base=not self.unshuffle)


class RTMoSRArch(Architecture[RTMoSR]):
    def __init__(self):
        rep = ('alpha', 'conv1.k0', 'conv1.b0', 'conv1.k1', 'conv1.b1', 'conv2.weight', 'conv2.bias', 'conv3.sk.weight', 'conv3.conv.0.weight',
               'conv3.conv.1.weight', 'conv3.conv.2.weight', 'conv3.eval_conv.weight', 'conv_3x3_rep.weight', 'conv_3x3_rep.bias')
        keys = ['body.0.norm.scale', 'body.0.norm.offset'] + [f'{p}.{k}' for p in ('body.0.fc1', 'body.0.conv.0.poll.1') for k in rep]
        keys += [f'body.0.conv.1.{k}' for k in ('alpha1', 'alpha2', 'alpha3', 'alpha4', 'conv1x1.weight', 'conv3x3.weight', 'conv5x5.weight',
                                                'conv5x5_reparam.weight', 'conv5x5_reparam.bias')]
        keys += [f'to_img.0.{k}' for k in rep[:11]]
        super().__init__(uid='RTMoSR', detect=KeyCondition.has_all(*keys))

    def load(self, state: Mapping[str, object]):
        # rtmosr/__init__.py:85-104, quirks included: the metadata's upscale is always 2, and for the unshuffle variant `scale` is
        # derived from the unshuffle factor (which coincides with the model's scale for the only supported pairing, 2 <-> 2)
        unshuffle = 'to_feat.1.alpha' in state
        if unshuffle:
            scale = math.isqrt(state['to_feat.1.conv_3x3_rep.weight'].shape[1] // 3)
            dim = state['to_feat.1.conv_3x3_rep.weight'].shape[0]
        else:
            scale = math.isqrt(state['to_img.0.conv_3x3_rep.weight'].shape[0] // 3)
            dim = state['to_feat.conv_3x3_rep.weight'].shape[0]
        dccm = 'body.0.fc2.alpha' in state
        se = 'body.0.conv.2.squeezing.0.weight' in state
        ffn = state['body.0.fc1.conv_3x3_rep.weight'].shape[0] / dim / 2
        n_blocks = get_seq_len(state, 'body')
        model = RTMoSR(scale=scale, dim=dim, ffn_expansion=ffn, n_blocks=n_blocks, unshuffle_mod=unshuffle, dccm=dccm, se=se)
        return self._enhance_model(model=model, in_channels=3, out_channels=3, upscale=int(2), name='RTMoSR')
