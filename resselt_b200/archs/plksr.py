"""RealPLKSR and the original PLKSR on the B200 engine (uid 'PLKSR': the reference registers both under one plugin).

Reference: /root/reference/resselt/archs/plksr/rplksr.py:96-147 (RealPLKSR), :52-93 (PLKBlock), :10-49 (DCCM, PLKConv2d, EA);
/root/reference/resselt/archs/plksr/plksr.py:311-377 (PLKSR), :259-308 (PLKBlock), :18-52 (CCM / ICCM / DCCM), :55-98 (PLKConv2d),
:101-117 (RectSparsePLKConv2d), :120-236 (SparsePLKConv2d), :239-248 (EA) and the loader
/root/reference/resselt/archs/plksr/__init__.py:10-121.  Only the PixelShuffle heads are in scope (``dysample=False``).

Notes that shape the plan:
  * the "partial large kernel" conv is a *dense* 17x17 conv on the first ``pdim`` channels, applied in place in eval
    mode (rplksr.py:36-37); here DCCM's second conv writes those channels to a side buffer, the 17x17 conv writes them
    back next to the untouched channels (a spatial conv cannot run in place on a tiled GPU kernel);
  * EA ``x * sigmoid(conv(x))`` is the conv's epilogue; GroupNorm + block skip is one two-pass kernel pair
    (a global reduction per sample and group: the only grid-wide dependency in the network);
  * ``feats(x) + repeat_interleave(x, r^2)`` followed by PixelShuffle is "add x[c][h][w] to every sub-pixel" — the same
    fused store Compact uses.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len, pixelshuffle_scale
from ._common import conv_specs, dysample_specs, emit_dysample


class RealPLKSR(EngineModule):
    def __init__(
        self,
        in_ch: int = 3,
        dim: int = 64,
        n_blocks: int = 28,
        upscaling_factor: int = 4,
        kernel_size: int = 17,
        split_ratio: float = 0.25,
        use_ea: bool = True,
        norm_groups: int = 4,
        dysample: bool = False,
        seed: int = 0,
    ):
        if dysample and upscaling_factor == 1:
            raise NotImplementedError('RealPLKSR 1x with DySample (no end convolution, rplksr.py:139) is not supported')
        pdim = int(dim * split_ratio)
        if pdim % 8 != 0 or dim % 8 != 0 or (dim // norm_groups) % 8 != 0:
            raise NotImplementedError('dim, dim*split_ratio and dim/norm_groups must be multiples of 8 (planar-8 layout)')
        r2 = upscaling_factor * upscaling_factor
        specs = conv_specs('feats.0', in_ch, dim, 3)
        for i in range(1, n_blocks + 1):
            p = f'feats.{i}'
            specs += conv_specs(f'{p}.channel_mixer.0', dim, dim * 2, 3)
            specs += conv_specs(f'{p}.channel_mixer.2', dim * 2, dim, 3)
            specs += conv_specs(f'{p}.lk.conv', pdim, pdim, kernel_size)
            if use_ea:
                specs += conv_specs(f'{p}.attn.f.0', dim, dim, 3)
            specs += conv_specs(f'{p}.refine', dim, dim, 1)
            specs += [(f'{p}.norm.weight', (dim,), 'affine_w'), (f'{p}.norm.bias', (dim,), f'bias:{dim}')]
        specs += conv_specs(f'feats.{n_blocks + 2}', dim, in_ch * r2, 3)  # feats.{n_blocks+1} is the (param-free) Dropout2d
        self_groups = in_ch if upscaling_factor % 2 != 0 else 4  # rplksr.py:134
        if dysample:
            specs += dysample_specs('to_img', in_ch * r2, in_ch, upscaling_factor, groups=self_groups)
        super().__init__(specs, in_ch, in_ch, upscaling_factor, seed=seed)
        self.dysample, self.dys_groups = dysample, self_groups
        self.dim, self.pdim, self.n_blocks, self.kernel_size = dim, pdim, n_blocks, kernel_size
        self.use_ea, self.norm_groups = use_ea, norm_groups

    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, pdim = self.dim, self.pdim
        xs = [pb.buffer(dim), pb.buffer(dim)]  # block input / output ping-pong
        mixed = pb.buffer(2 * dim)             # DCCM hidden (and, afterwards, scratch for EA / refine outputs)
        lk_in = pb.buffer(pdim)
        t = pb.buffer(dim)                     # [lk(t[:pdim]) | t[pdim:]]
        pb.conv(INPUT, xs[0], w['feats.0.weight'], w['feats.0.bias'])
        for i in range(1, self.n_blocks + 1):
            p = f'feats.{i}'
            x_in, x_out = xs[(i - 1) % 2], xs[i % 2]
            pb.conv(x_in, mixed, w[f'{p}.channel_mixer.0.weight'], w[f'{p}.channel_mixer.0.bias'], act=N.ACT_MISH)
            pb.conv(mixed, lk_in, w[f'{p}.channel_mixer.2.weight'], w[f'{p}.channel_mixer.2.bias'], dst2=t.slice(pdim, dim - pdim))
            pb.conv(lk_in, t.slice(0, pdim), w[f'{p}.lk.conv.weight'], w[f'{p}.lk.conv.bias'])
            cur = t
            if self.use_ea:
                ea = mixed.slice(0, dim)
                pb.conv(t, ea, w[f'{p}.attn.f.0.weight'], w[f'{p}.attn.f.0.bias'], act=N.ACT_SIGMOID, combine=N.COMB_MUL, res1=t)
                cur = ea
            refined = mixed.slice(dim, dim)
            pb.conv(cur, refined, w[f'{p}.refine.weight'], w[f'{p}.refine.bias'])
            pb.groupnorm(refined, x_out, self.norm_groups, w[f'{p}.norm.weight'], w[f'{p}.norm.bias'], eps=1e-5, skip=x_in)
        last = f'feats.{self.n_blocks + 2}'
        if not self.dysample:
            pb.conv(xs[self.n_blocks % 2], OUTPUT, w[f'{last}.weight'], w[f'{last}.bias'], ps=self.upscale, add_base=True)
            return
        # DySample head (rplksr.py:133-147): it reads feats(x) + repeat_interleave(x, r^2) as a 3 r^2-channel low-res map.  The
        # repeated input is a 1-tap conv of the caller's tensor (channel c r^2 + k copies input channel c); the last conv adds it.
        r2, cin = self.upscale * self.upscale, self.in_channels
        rep_w = torch.zeros(cin * r2, cin, 3, 3, dtype=torch.float64)
        for c in range(cin):
            rep_w[c * r2:(c + 1) * r2, c, 1, 1] = 1.0
        rep, pre = pb.buffer(cin * r2), pb.buffer(cin * r2)
        pb.conv(INPUT, rep, rep_w, None)
        pb.conv(xs[self.n_blocks % 2], pre, w[f'{last}.weight'], w[f'{last}.bias'], combine=N.COMB_AXPY, res1=rep)
        emit_dysample(pb, w, 'to_img', pre, self.out_channels, self.upscale, groups=self.dys_groups)


def _dense_lk_kernel(w, p: str, lk_type: str, pdim: int, kmax: int, dilations: Sequence[int], with_idt: bool):
    """The partial large-kernel layer of the original PLKSR as ONE dense kmax x kmax conv (weight, bias), in fp64.

    'PLK' is already that (plksr.py:55-81).  'SparsePLK' sums dilated k x k convs (plksr.py:153-164): a conv with dilation d is
    a dense ((k-1)d+1)-kernel with zeros between the taps, every branch is 'same'-padded and centred, so the sum of the
    zero-padded kernels is exact (this is the model's own ``convert``, plksr.py:206-236).  'RectSparsePLK' sums an m x n, an
    n x m and an n x n conv (plksr.py:101-117) — again centred sub-kernels of one m x m kernel.  ``with_idt`` adds the
    identity at the centre tap."""
    k = torch.zeros(pdim, pdim, kmax, kmax, dtype=torch.float64)
    b = torch.zeros(pdim, dtype=torch.float64)

    def add(weight, bias, dil=1):
        kh, kw = weight.shape[2], weight.shape[3]
        eh, ew = (kh - 1) * dil + 1, (kw - 1) * dil + 1
        dense = torch.zeros(pdim, pdim, eh, ew, dtype=torch.float64)
        dense[:, :, ::dil, ::dil] = weight
        ph, pw = (kmax - eh) // 2, (kmax - ew) // 2
        k[:, :, ph:ph + eh, pw:pw + ew] += dense
        b.add_(bias)

    if lk_type == 'PLK':
        add(w[f'{p}.conv.weight'], w[f'{p}.conv.bias'])
    elif lk_type == 'SparsePLK':
        n_conv = 0
        while f'{p}.convs.{n_conv}.weight' in w:
            n_conv += 1
        for i in range(n_conv):
            add(w[f'{p}.convs.{i}.weight'], w[f'{p}.convs.{i}.bias'], dilations[i] if i < len(dilations) else 1)
    else:
        for name in ('mn_conv', 'nm_conv', 'nn_conv'):
            add(w[f'{p}.{name}.weight'], w[f'{p}.{name}.bias'])
    if with_idt:
        k[:, :, kmax // 2, kmax // 2] += torch.eye(pdim, dtype=torch.float64)
    return k, b


class PLKSR(EngineModule):
    """The original PLKSR (reference class ``plksr``): CCM / ICCM / DCCM mixer with GELU, partial large-kernel layer of any of the
    three types, optional element-wise attention, 1x1 refine + block skip; RGB only (plksr.py:329, 347)."""

    _MIXER_K = {'CCM': (3, 1), 'ICCM': (1, 3), 'DCCM': (3, 3)}

    def __init__(
        self,
        dim: int = 64,
        n_blocks: int = 28,
        upscaling_factor: int = 4,
        ccm_type: str = 'DCCM',
        kernel_size: int = 17,
        split_ratio: float = 0.25,
        lk_type: str = 'PLK',
        use_max_kernel: bool = False,
        sparse_kernels: Sequence[int] = (5, 5, 5, 5),
        sparse_dilations: Sequence[int] = (1, 2, 3, 4),
        with_idt: bool = False,
        use_ea: bool = True,
        seed: int = 0,
    ):
        if ccm_type not in self._MIXER_K:
            raise ValueError(f'Unknown CCM type: {ccm_type}')
        if lk_type not in ('PLK', 'SparsePLK', 'RectSparsePLK'):
            raise ValueError(f'Unknown LK type: {lk_type}')
        pdim = int(dim * split_ratio)
        if pdim % 8 != 0 or dim % 8 != 0:
            raise NotImplementedError('dim and dim*split_ratio must be multiples of 8 (planar-8 layout)')
        kmax = kernel_size
        if lk_type == 'SparsePLK':
            kmax = max([kernel_size] + [(k - 1) * d + 1 for k, d in zip(sparse_kernels, sparse_dilations)])  # plksr.py:135-138
        if kmax % 2 == 0 or (lk_type == 'RectSparsePLK' and (kernel_size // 3) % 2 == 0):
            raise NotImplementedError('large-kernel sizes must be odd')
        k0, k2 = self._MIXER_K[ccm_type]
        r2 = upscaling_factor * upscaling_factor
        specs = conv_specs('feats.0', 3, dim, 3)
        for i in range(1, n_blocks + 1):
            p = f'feats.{i}'
            specs += conv_specs(f'{p}.channe_mixer.0', dim, dim * 2, k0)  # sic: the reference attribute is misspelt (plksr.py:283)
            specs += conv_specs(f'{p}.channe_mixer.2', dim * 2, dim, k2)
            if lk_type == 'PLK':
                specs += conv_specs(f'{p}.lk.conv', pdim, pdim, kernel_size)
            elif lk_type == 'SparsePLK':
                for j, (k, _) in enumerate(zip(sparse_kernels, sparse_dilations)):
                    specs += conv_specs(f'{p}.lk.convs.{j}', pdim, pdim, k)
                if use_max_kernel:
                    specs += conv_specs(f'{p}.lk.convs.{len(list(zip(sparse_kernels, sparse_dilations)))}', pdim, pdim, kmax)
            else:
                m, n = kernel_size, kernel_size // 3
                for name, shape in (('mn_conv', (m, n)), ('nm_conv', (n, m)), ('nn_conv', (n, n))):
                    specs += [(f'{p}.lk.{name}.weight', (pdim, pdim, *shape), 'conv_w'), (f'{p}.lk.{name}.bias', (pdim,), f'bias:{pdim * shape[0] * shape[1]}')]
            if use_ea:
                specs += conv_specs(f'{p}.attn.f.0', dim, dim, 3)
            specs += conv_specs(f'{p}.refine', dim, dim, 1)
        specs += conv_specs(f'feats.{n_blocks + 1}', dim, 3 * r2, 3)
        super().__init__(specs, 3, 3, upscaling_factor, seed=seed)
        self.dim, self.pdim, self.n_blocks, self.kernel_size, self.kmax = dim, pdim, n_blocks, kernel_size, kmax
        self.ccm_type, self.lk_type, self.use_ea, self.with_idt = ccm_type, lk_type, use_ea, with_idt
        self.sparse_dilations = list(sparse_dilations)

    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, pdim = self.dim, self.pdim
        xs = [pb.buffer(dim), pb.buffer(dim)]
        mixed = pb.buffer(2 * dim)
        lk_in = pb.buffer(pdim)
        t = pb.buffer(dim)
        pb.conv(INPUT, xs[0], w['feats.0.weight'], w['feats.0.bias'])
        for i in range(1, self.n_blocks + 1):
            p = f'feats.{i}'
            x_in, x_out = xs[(i - 1) % 2], xs[i % 2]
            pb.conv(x_in, mixed, w[f'{p}.channe_mixer.0.weight'], w[f'{p}.channe_mixer.0.bias'], act=N.ACT_GELU)
            pb.conv(mixed, lk_in, w[f'{p}.channe_mixer.2.weight'], w[f'{p}.channe_mixer.2.bias'], dst2=t.slice(pdim, dim - pdim))
            lk_w, lk_b = _dense_lk_kernel(w, f'{p}.lk', self.lk_type, pdim, self.kmax, self.sparse_dilations, self.with_idt)
            pb.conv(lk_in, t.slice(0, pdim), lk_w, lk_b)
            cur = t
            if self.use_ea:
                ea = mixed.slice(0, dim)
                pb.conv(t, ea, w[f'{p}.attn.f.0.weight'], w[f'{p}.attn.f.0.bias'], act=N.ACT_SIGMOID, combine=N.COMB_MUL, res1=t)
                cur = ea
            pb.conv(cur, x_out, w[f'{p}.refine.weight'], w[f'{p}.refine.bias'], combine=N.COMB_AXPY, res1=x_in)  # refine(x) + x_skip
        last = f'feats.{self.n_blocks + 1}'
        pb.conv(xs[self.n_blocks % 2], OUTPUT, w[f'{last}.weight'], w[f'{last}.bias'], ps=self.upscale, add_base=True)


class PLKSRArch(Architecture[RealPLKSR]):
    def __init__(self):
        super().__init__(
            uid='PLKSR',
            detect=KeyCondition.has_all(
                'feats.0.weight',
                KeyCondition.has_any('feats.1.lk.conv.weight', 'feats.1.lk.convs.0.weight', 'feats.1.lk.mn_conv.weight'),
                'feats.1.refine.weight',
                KeyCondition.has_any('feats.1.channe_mixer.0.weight', 'feats.1.channel_mixer.0.weight'),
            ),
        )

    def load(self, state_dict: Mapping[str, object]):
        if 'feats.1.channel_mixer.0.weight' not in state_dict:
            return self._load_plksr(state_dict)  # the typo'd 'channe_mixer' marks the original PLKSR (plksr/__init__.py:41-42)
        in_nc = state_dict['feats.0.weight'].shape[1]
        dim = state_dict['feats.0.weight'].shape[0]
        total = get_seq_len(state_dict, 'feats')
        scale = pixelshuffle_scale(state_dict[f'feats.{total - 1}.weight'].shape[0], in_nc)
        lk = state_dict['feats.1.lk.conv.weight']
        model = RealPLKSR(
            in_ch=in_nc,
            dim=dim,
            n_blocks=total - 3,
            upscaling_factor=scale,
            kernel_size=lk.shape[2],
            split_ratio=lk.shape[0] / dim,
            use_ea='feats.1.attn.f.0.weight' in state_dict,
            norm_groups=4,  # not recoverable from the checkpoint (plksr/__init__.py:116)
            dysample='to_img.init_pos' in state_dict,
        )
        return self._enhance_model(model=model, in_channels=in_nc, out_channels=in_nc, upscale=scale, name='RealPLKSR')

    def _load_plksr(self, state_dict: Mapping[str, object]):
        """Hyper-parameter inference of the original PLKSR, rule for rule as plksr/__init__.py:30-96 (sparse kernel sizes and
        dilations, ``use_max_kernel`` and ``with_idt`` are not recoverable there either: defaults)."""
        in_nc = state_dict['feats.0.weight'].shape[1]
        dim = state_dict['feats.0.weight'].shape[0]
        total = get_seq_len(state_dict, 'feats')
        scale = pixelshuffle_scale(state_dict[f'feats.{total - 1}.weight'].shape[0], in_nc)
        shapes = (state_dict['feats.1.channe_mixer.0.weight'].shape[2], state_dict['feats.1.channe_mixer.2.weight'].shape[2])
        ccm = {(3, 1): 'CCM', (3, 3): 'DCCM', (1, 3): 'ICCM'}.get(shapes)
        if ccm is None:
            raise ValueError('Unknown CCM type')
        kernel_size = 17
        if 'feats.1.lk.conv.weight' in state_dict:
            lk_type, lk = 'PLK', state_dict['feats.1.lk.conv.weight']
            kernel_size = lk.shape[2]
        elif 'feats.1.lk.convs.0.weight' in state_dict:
            lk_type, lk = 'SparsePLK', state_dict['feats.1.lk.convs.0.weight']
        elif 'feats.1.lk.mn_conv.weight' in state_dict:
            lk_type, lk = 'RectSparsePLK', state_dict['feats.1.lk.mn_conv.weight']
            kernel_size = lk.shape[2]
        else:
            raise ValueError('Unknown LK type')
        model = PLKSR(dim=dim, n_blocks=total - 2, upscaling_factor=scale, ccm_type=ccm, kernel_size=kernel_size, split_ratio=lk.shape[0] / dim,
                      lk_type=lk_type, use_ea='feats.1.attn.f.0.weight' in state_dict)
        return self._enhance_model(model=model, in_channels=in_nc, out_channels=in_nc, upscale=scale, name='PLKSR')
