"""SPAN++ (SpanPP) on the B200 engine — SURVEY.md section 8f rank 2, the first of the SPAN descendants.

Reference: /root/reference/resselt/archs/spanpp/arch.py:315-373 (model), :195-216 (SPAB: same block as SPAN's, built from
``RepConv``), :152-193 (RepConv: three parallel 3x3 branches ``alpha[0] * SeqConv3x3 + alpha[1] * Conv2d + alpha[2] * Conv3XC`` that
``.eval()`` fuses into one 3x3), :244-312 (IGConv: the up-sampling conv whose kernel is generated per scale by a small implicit
network over Fourier features, pre-computed in ``train()``), loader /root/reference/resselt/archs/spanpp/__init__.py:8-132.

Lowering: every RepConv is merged once per plan on the host in fp64 (the checkpoint's ``conv_3x3_rep.*`` / ``eval_conv.*`` are
dead values exactly as in SPAN: ``.eval()`` overwrites them from the branches); the IGConv kernel of the requested scale is
evaluated once per plan on the host (it depends on weights only); the graph is SPAN's — same kernels, same epilogues, no input
normalisation, no bias on the last conv.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, ParamSpec, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len
from ._common import conv3xc_specs, conv_specs, merge_conv3xc, merge_pointwise_into_conv


def repconv_specs(prefix: str, cin: int, cout: int) -> List[ParamSpec]:
    """Parameter names of a RepConv (spanpp/arch.py:152-166): SeqConv3x3 (k0, b0, k1, b1 with depth multiplier 2), a plain 3x3,
    a Conv3XC, the (dead) fused ``conv_3x3_rep`` and the three mixing weights ``alpha``."""
    mid = 2 * cout
    return (
        [(f'{prefix}.alpha', (3,), 'affine_w')]
        + [(f'{prefix}.conv1.k0', (mid, cin, 1, 1), 'conv_w'), (f'{prefix}.conv1.b0', (mid,), f'bias:{cin}')]
        + [(f'{prefix}.conv1.k1', (cout, mid, 3, 3), 'conv_w'), (f'{prefix}.conv1.b1', (cout,), f'bias:{mid * 9}')]
        + conv_specs(f'{prefix}.conv2', cin, cout, 3)
        + conv3xc_specs(f'{prefix}.conv3', cin, cout)
        + conv_specs(f'{prefix}.conv_3x3_rep', cin, cout, 3)
    )


def merge_repconv(w: Dict[str, torch.Tensor], prefix: str):
    """RepConv.fuse (spanpp/arch.py:164-173) in closed form, fp64:
    SeqConv3x3.rep_params (:136-149): K[o,i] = sum_m k1[o,m] k0[m,i], B[o] = sum_{m,taps} k1[o,m,tap] b0[m] + b1[o] (the training
    branch pads the 1x1 output with its bias, which is what makes this exact for zero padding of the input);
    Conv3XC.update_params (:61-97) as for SPAN; result = alpha[0] * SeqConv + alpha[1] * conv2 + alpha[2] * Conv3XC."""
    f64 = torch.float64
    g = lambda k: w[f'{prefix}.{k}'].to(f64)
    k0, b0, k1, b1 = g('conv1.k0')[:, :, 0, 0], g('conv1.b0'), g('conv1.k1'), g('conv1.b1')
    w1 = torch.einsum('omhw,mi->oihw', k1, k0)
    bias1 = torch.einsum('omhw,m->o', k1, b0) + b1
    w3, bias3 = merge_conv3xc(w, f'{prefix}.conv3')
    a = g('alpha')
    return a[0] * w1 + a[1] * g('conv2.weight') + a[2] * w3, a[0] * bias1 + a[1] * g('conv2.bias') + a[2] * bias3


def igconv_kernel(w: Dict[str, torch.Tensor], prefix: str, dim: int, ksize: int, scale: int, max_scale: int) -> torch.Tensor:
    """The conv kernel IGConv generates for ``scale`` (spanpp/arch.py:289-312), [3 * scale^2, dim, k, k], fp64.

    For every (input channel, tap) a latent code of ``implicit_dim`` frequencies is turned into Fourier features of the s x s
    sub-pixel centre coordinates (in [-1, 1], x then y) plus a learned phase of the cell size 2 / min(scale, max_scale), scaled
    by the amplitude code and decoded by the 1x1-conv MLP ``query_kernel`` into the RGB kernel value of each sub-pixel."""
    f64 = torch.float64
    g = lambda k: w[f'{prefix}.{k}'].to(f64)
    s = int(scale)
    centres = -1.0 + 1.0 / s + (2.0 / s) * torch.arange(s, dtype=f64)            # make_coord (:219-231)
    yy, xx = torch.meshgrid(centres, centres, indexing='ij')
    freq, amp = g('freq'), g('amplitude')                                          # [dim * k * k, implicit_dim, 1, 1]
    half = freq.shape[1] // 2
    arg = freq[:, :half] * xx + freq[:, half:] * yy                                # [K, D/2, s, s]
    cell = torch.full((1, 1, s, s), 2.0 / min(s, int(max_scale)), dtype=f64)
    arg = arg + F.conv2d(cell, g('phase.weight'), g('phase.bias'))
    t = torch.cat([torch.cos(math.pi * arg), torch.sin(math.pi * arg)], dim=1) * amp
    n_layers = get_seq_len(w, f'{prefix}.query_kernel')
    for i in range(0, n_layers, 2):                                                # conv1x1, ReLU, ..., conv1x1 -> 3
        t = F.conv2d(t, g(f'query_kernel.{i}.weight'), g(f'query_kernel.{i}.bias'))
        if i + 1 < n_layers:
            t = F.relu(t)
    # '(Cin Kh Kw) RGB rh rw -> (RGB rh rw) Cin Kh Kw'
    return t.reshape(dim, ksize, ksize, 3, s, s).permute(3, 4, 5, 0, 1, 2).reshape(3 * s * s, dim, ksize, ksize)


class SpanPP(EngineModule):
    def __init__(
        self,
        *,
        num_in_ch: int = 3,
        feature_channels: int = 48,
        scale_list=(1, 2, 3, 4),
        eval_base_scale: int = 2,
        ig_kernel_size: int = 3,
        implicit_dim: int = 256,
        latent_layers: int = 4,
        seed: int = 0,
        **kwargs,  # the reference's constructor swallows unknown keywords too (its loader passes ``ig_kernel=``, :120-129)
    ):
        f = feature_channels
        if f % 8 != 0:
            raise ValueError('feature_channels must be a multiple of 8 for the planar-8 activation layout')
        if implicit_dim % 2 != 0:
            raise AssertionError('implicit_dim must be even')  # IGConv.__init__ (:252)
        scales = sorted(set(int(s) for s in scale_list))
        specs = repconv_specs('conv0', num_in_ch, f)
        for i in range(1, 7):
            for c in ('c1_r', 'c2_r', 'c3_r'):
                specs += repconv_specs(f'block_{i}.{c}', f, f)
        specs += conv_specs('conv_cat', 4 * f, f, 1)
        specs += repconv_specs('conv_2', f, f)
        k2 = f * ig_kernel_size * ig_kernel_size
        specs += conv_specs('upsampler.phase', 1, implicit_dim // 2, 1)
        specs += [('upsampler.freq', (k2, implicit_dim, 1, 1), 'normal:0.5'), ('upsampler.amplitude', (k2, implicit_dim, 1, 1), 'normal:0.3')]
        for i in range(latent_layers):
            specs += conv_specs(f'upsampler.query_kernel.{2 * i}', implicit_dim, implicit_dim, 1, gain=2.0)
        specs += conv_specs(f'upsampler.query_kernel.{2 * latent_layers}', implicit_dim, 3, 1, gain=2.0)
        specs += [('MetaIGConv', torch.tensor(scales, dtype=torch.uint8), 'buffer_tensor')]
        super().__init__(specs, num_in_ch, 3, eval_base_scale, seed=seed)
        self.feature_channels, self.scale_list, self.base_scale = f, scales, int(eval_base_scale)
        self.ig_kernel_size, self.max_scale = int(ig_kernel_size), max(scales)
        self._scale = self.base_scale

    def load_state_dict(self, state_dict, *args, **kwargs):
        # like the reference (spanpp/arch.py:354-356): the scale list is the module's own, whatever the checkpoint says
        state_dict = dict(state_dict)
        state_dict['MetaIGConv'] = self.MetaIGConv
        return super().load_state_dict(state_dict, *args, **kwargs)

    # one native plan per (device, dtype, scale): the IGConv kernel and the output geometry depend on the scale
    def _plan_variant(self):
        return self._scale

    def _plan_io_for_variant(self):
        return (self.in_channels, 3, self._scale)

    def forward(self, x: torch.Tensor, scale: Optional[int] = None) -> torch.Tensor:
        """``scale=None`` selects ``eval_base_scale`` like the reference (:289-291); any scale of ``scale_list`` may be asked for."""
        s = self.base_scale if scale is None else int(scale)
        if s not in self.scale_list:
            raise KeyError(str(s))  # the reference looks the pre-computed kernel up in a dict keyed by str(scale) (:296)
        self._scale = s
        try:
            return super().forward(x)
        finally:
            self._scale = self.base_scale

    @property
    def receptive_radius(self) -> int:
        return 1 + 6 * 3 + 1 + self.ig_kernel_size // 2

    def build_plan(self, pb: PlanBuilder, w) -> None:
        f = self.feature_channels
        cat = pb.buffer(4 * f)  # [conv0 out | conv_2 out | block_1 out | act(block_6.c1_r)] == the reference's torch.cat (:366)
        feat, tail, b1, o1_end = (cat.slice(i * f, f) for i in range(4))
        t1, t2, p0, p1 = (pb.buffer(f) for _ in range(4))

        def spab(prefix, src, dst, o1):
            pb.conv(src, o1, *merge_repconv(w, f'{prefix}.c1_r'), act=N.ACT_SILU)
            pb.conv(o1, t2, *merge_repconv(w, f'{prefix}.c2_r'), act=N.ACT_SILU)
            pb.conv(t2, dst, *merge_repconv(w, f'{prefix}.c3_r'), combine=N.COMB_SPAB_GATE, res1=src)

        pb.conv(INPUT, feat, *merge_repconv(w, 'conv0'))
        spab('block_1', feat, b1, t1)
        spab('block_2', b1, p0, t1)
        spab('block_3', p0, p1, t1)
        spab('block_4', p1, p0, t1)
        spab('block_5', p0, p1, t1)
        spab('block_6', p1, p0, o1_end)
        pb.conv(p0, tail, *merge_repconv(w, 'conv_2'))
        s = pb.upscale
        kernel = igconv_kernel(w, 'upsampler', f, self.ig_kernel_size, s, self.max_scale)
        if self.ig_kernel_size == 3:
            # conv_cat (1x1) merged into the IGConv kernel: one 3x3 conv over the concat (see span.py)
            wm, bm, border = merge_pointwise_into_conv(w['conv_cat.weight'], w['conv_cat.bias'], kernel, None)
            pb.conv(cat, OUTPUT, wm, bm, ps=s, border_bias=border)
        else:
            pb.conv(cat, t1, w['conv_cat.weight'], w['conv_cat.bias'])
            pb.conv(t1, OUTPUT, kernel, None, ps=s)


class SpanPPArch(Architecture[SpanPP]):
    def __init__(self):
        # the reference lists ~100 keys of conv0 / block_1 / block_2 (spanpp/__init__.py:11-113); these identify the same layout
        parts = ('alpha', 'conv1.k0', 'conv1.b0', 'conv1.k1', 'conv1.b1', 'conv2.weight', 'conv2.bias', 'conv3.sk.weight', 'conv3.sk.bias',
                 'conv3.conv.0.weight', 'conv3.conv.1.weight', 'conv3.conv.2.weight', 'conv3.eval_conv.weight', 'conv_3x3_rep.weight', 'conv_3x3_rep.bias')
        keys = [f'{p}.{k}' for p in ('conv0', 'block_1.c1_r', 'block_1.c2_r', 'block_1.c3_r', 'block_2.c1_r') for k in parts]
        super().__init__(uid='SpanPP', detect=KeyCondition.has_all(*keys))

    def load(self, state_dict: Mapping[str, object]):
        dim, in_ch = state_dict['conv0.conv_3x3_rep.weight'].shape[:2]
        scales = state_dict['MetaIGConv'].tolist() if 'MetaIGConv' in state_dict else [1, 2, 3, 4]
        _, implicit_dim = state_dict['upsampler.freq'].shape[:2]
        latent_layers = get_seq_len(state_dict, 'upsampler.query_kernel') // 2
        # NB the reference derives ig_kernel_size from upsampler.freq but passes it under a name its constructor ignores
        # (``ig_kernel=``, spanpp/__init__.py:120-129), so the model is always built with the default 3x3 kernel; same here
        model = SpanPP(num_in_ch=in_ch, feature_channels=dim, scale_list=scales, eval_base_scale=2, implicit_dim=implicit_dim,
                       latent_layers=latent_layers)
        # the reference hands the scale LIST to the metadata's upscale field (:132); kept for drop-in behaviour
        return self._enhance_model(model=model, in_channels=in_ch, out_channels=in_ch, upscale=scales, name='SpanPP')
