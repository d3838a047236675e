"""RealPLKSR on the B200 engine (uid 'PLKSR', the reference registers PLKSR and RealPLKSR under one plugin).

Reference: /root/reference/resselt/archs/plksr/rplksr.py:96-147 (model), :52-93 (PLKBlock), :10-49 (DCCM, PLKConv2d, EA)
and the loader /root/reference/resselt/archs/plksr/__init__.py:10-121.  Only the PixelShuffle head
(``dysample=False``) is in scope; the original PLKSR (plksr.py) is a "next" row.

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

from typing import Mapping

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len, pixelshuffle_scale
from ._common import conv_specs


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
        if dysample:
            raise NotImplementedError('RealPLKSR with the DySample head is not supported by the B200 engine (see DESIGN.md)')
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
        super().__init__(specs, in_ch, in_ch, upscaling_factor, seed=seed)
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
            # the typo'd 'channe_mixer' marks the original PLKSR (plksr/__init__.py:41-42)
            raise NotImplementedError('original PLKSR checkpoints are not supported yet (RealPLKSR only, see DESIGN.md)')
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
