"""GateRV3 on the B200 engine — SURVEY.md section 8f rank 2, the third of the SPAN descendants.

Reference: /root/reference/resselt/archs/gaterv3/arch.py:695-802 (model: a SPAN branch at full resolution next to a four-level
gated-CNN U-Net), :487-508 (SPAB), :376-484 (Conv3XC), :640-667 (MetaGated: RMSNorm -> 1x1 -> grouped 3x3 -> SimpleGate ->
simplified channel attention -> ``* gamma0 + short`` -> ``GatedCNNBlock * gamma1 + x``), :594-629 (GatedCNNBlock), :527-557
(InceptionDWConv2d), :511-524 (RMSNorm), :670-692 (Down / Upsample / Block), :241-373 (UniUpsampleV3),
loader /root/reference/resselt/archs/gaterv3/__init__.py:8-157.

Lowering.  One plan with base divisor 16 holds all five grids (H, H/2, H/4, H/8, H/16).
  * Conv3XC (bias-free in the SPABs) is merged on the host in fp64; the SPAB gate is the third conv's epilogue; the four maps of
    ``torch.cat([x, sisr, sisr_short, sisr_out])`` are channel ranges of one buffer, and ``sisr_cat_conv`` runs LAST with the decoder
    output as its residual (``x + sisr``).
  * MetaGated: the grouped 3x3 conv (two channels per group) runs as block-diagonal dense convs on the tensor cores, one per
    <= 64-channel chunk and half, the second half first so that SimpleGate (x1 * x2) is the first half's epilogue (COMB_MUL);
    ``x * sca(x) * gamma0 + short`` is one op (deterministic global mean, 1x1 conv on the mean, scale + shortcut: rsb_op_kind 10);
    ``glob(x) * gamma1 + x`` is a per-channel affine op (kind 11).
  * GatedCNNBlock: fc1 as three convs (identity part, conv part, gate part) so that the gate conv's epilogue does
    ``mish(g) * cat(i, c)``; the inception depthwise conv (identity | 3x3 | 1x11 | 11x1 channel groups) is ONE depthwise 11x11 kernel
    built on the host, whose zero taps the kernel skips per 8-channel plane; fc2's epilogue is the Mish.
  * Down = conv + PixelUnshuffle(2) (the unshuffle op of RTMoSR), Upsample = conv + PixelShuffle(2) (plain shuffle op); the
    encoder's level outputs are written straight into the second half of the decoder's concat buffers.
  * Heads: ``pixelshuffle`` (default), ``pixelshuffledirect``, ``nearest+conv`` and ``pa_up`` (2^n and x3; the nearest upsample folded into 2x2
    phase convs), ``transpose+conv`` (ConvTranspose2d as sub-pixel phase convs), ``dysample`` with a 1x1 end conv, and the plain conv of
    scale 1.  ``lda`` (deformable LDA-AQU) is refused at load.  The latent self-attention variant (``attention=True``) maps onto DAT's channel-attention op behind a 1x1
    conv and a depthwise 3x3.
Reflect padding to a multiple of 16, the crop and ``+ gamma * nearest(inp)`` are host glue around the plan (arch.py:789-802).
"""
from __future__ import annotations

import math
from typing import List, Mapping, Sequence

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, ParamSpec, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len
from ._common import conv_specs, dysample_specs, emit_dysample
from .esrgan import upconv_phase_kernels

SAMPLE_MODS = ('conv', 'pixelshuffledirect', 'pixelshuffle', 'nearest+conv', 'dysample', 'transpose+conv', 'lda', 'pa_up')
SUPPORTED_MODS = ('pixelshuffledirect', 'pixelshuffle', 'nearest+conv', 'dysample', 'transpose+conv', 'pa_up')


def _conv3xc_specs(prefix: str, cin: int, cout: int, gain: int, bias: bool) -> List[ParamSpec]:
    def c(name, i, o, k):
        s = [(f'{prefix}.{name}.weight', (o, i, k, k), 'conv_w')]
        return s + ([(f'{prefix}.{name}.bias', (o,), f'bias:{i * k * k}')] if bias else [])
    return c('sk', cin, cout, 1) + c('conv.0', cin, cin * gain, 1) + c('conv.1', cin * gain, cout * gain, 3) + c('conv.2', cout * gain, cout, 1) + c('eval_conv', cin, cout, 3)


def merge_conv3xc_any(w, prefix: str):
    """Conv3XC.update_params (gaterv3/arch.py:432-463) in closed form, fp64, with or without biases: 1x1 -> 3x3 -> 1x1 plus the 1x1
    skip as one 3x3 (weight, bias or None).  ``eval_conv.*`` of the checkpoint is dead: the reference overwrites it."""
    f64 = torch.float64
    w1, w2, w3 = w[f'{prefix}.conv.0.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.conv.1.weight'].to(f64), w[f'{prefix}.conv.2.weight'].to(f64)[:, :, 0, 0]
    merged = torch.einsum('on,nmhw,mi->oihw', w3, w2, w1)
    merged[:, :, 1, 1] += w[f'{prefix}.sk.weight'].to(f64)[:, :, 0, 0]
    if f'{prefix}.conv.0.bias' not in w:
        return merged, None
    b1, b2, b3 = (w[f'{prefix}.conv.{i}.bias'].to(f64) for i in range(3))
    bias = w3 @ (torch.einsum('nmhw,m->n', w2, b1) + b2) + b3 + w[f'{prefix}.sk.bias'].to(f64)
    return merged, bias


def _spab_specs(prefix: str, dim: int) -> List[ParamSpec]:
    return sum((_conv3xc_specs(f'{prefix}.{c}', dim, dim, 2, False) for c in ('c1_r', 'c2_r', 'c3_r')), [])


ATT_HEADS = 16  # Attention(conv_channels, 16) in GatedCNNBlock (arch.py:616)


def _gated_cnn_specs(p: str, dim: int, att: bool = False) -> List[ParamSpec]:
    hidden, gc = int(1.5 * dim), int(dim * 0.125)
    specs: List[ParamSpec] = [(f'{p}.norm.scale', (dim,), 'affine_w'), (f'{p}.norm.offset', (dim,), 'normal:0.1')]
    specs += conv_specs(f'{p}.fc1', dim, 2 * hidden, 1)
    if att:  # Attention (arch.py:560-591): temperature, qkv 1x1 (no bias), depthwise 3x3 on 3 dim channels, project_out 1x1 (no bias)
        specs += [(f'{p}.token_mix.temperature', (ATT_HEADS, 1, 1), 'affine_w'), (f'{p}.token_mix.qkv.weight', (3 * dim, dim, 1, 1), 'conv_w*2.0'),
                  (f'{p}.token_mix.qkv_dwconv.weight', (3 * dim, 1, 3, 3), 'conv_w*2.0'), (f'{p}.token_mix.qkv_dwconv.bias', (3 * dim,), 'bias:9'),
                  (f'{p}.token_mix.project_out.weight', (dim, dim, 1, 1), 'conv_w')]
        return specs + conv_specs(f'{p}.fc2', hidden, dim, 1)
    specs += [(f'{p}.token_mix.dwconv_hw.weight', (gc, 1, 3, 3), 'conv_w*2.0'), (f'{p}.token_mix.dwconv_hw.bias', (gc,), 'bias:9'),
              (f'{p}.token_mix.dwconv_w.weight', (gc, 1, 1, 11), 'conv_w*2.0'), (f'{p}.token_mix.dwconv_w.bias', (gc,), 'bias:11'),
              (f'{p}.token_mix.dwconv_h.weight', (gc, 1, 11, 1), 'conv_w*2.0'), (f'{p}.token_mix.dwconv_h.bias', (gc,), 'bias:11')]
    return specs + conv_specs(f'{p}.fc2', hidden, dim, 1)


def _meta_gated_specs(p: str, dim: int) -> List[ParamSpec]:
    specs: List[ParamSpec] = [(f'{p}.gamma0', (1, dim, 1, 1), 'affine_w'), (f'{p}.gamma1', (1, dim, 1, 1), 'affine_w'),
                              (f'{p}.local.0.scale', (dim,), 'affine_w'), (f'{p}.local.0.offset', (dim,), 'normal:0.1')]
    specs += conv_specs(f'{p}.local.1', dim, 2 * dim, 1)
    specs += [(f'{p}.local.2.weight', (2 * dim, 2, 3, 3), 'conv_w*2.0'), (f'{p}.local.2.bias', (2 * dim,), 'bias:18')]
    specs += conv_specs(f'{p}.sca.1', dim, dim, 1)
    return specs + _gated_cnn_specs(f'{p}.glob', dim)


def merge_inception(w, p: str, dim: int):
    """InceptionDWConv2d (gaterv3/arch.py:527-557) as one depthwise 11x11 kernel [dim][1][11][11] + bias: the first dim - 3 gc channels
    pass through (centre tap 1), then gc channels each of a 3x3, a 1x11 and an 11x1 depthwise conv."""
    gc = int(dim * 0.125)
    k = torch.zeros(dim, 1, 11, 11, dtype=torch.float64)
    b = torch.zeros(dim, dtype=torch.float64)
    i0 = dim - 3 * gc
    k[:i0, 0, 5, 5] = 1.0
    k[i0:i0 + gc, :, 4:7, 4:7] = w[f'{p}.dwconv_hw.weight'].double()
    k[i0 + gc:i0 + 2 * gc, :, 5:6, :] = w[f'{p}.dwconv_w.weight'].double()
    k[i0 + 2 * gc:, :, :, 5:6] = w[f'{p}.dwconv_h.weight'].double()
    b[i0:i0 + gc], b[i0 + gc:i0 + 2 * gc], b[i0 + 2 * gc:] = (w[f'{p}.dwconv_{n}.bias'].double() for n in ('hw', 'w', 'h'))
    return k, b


def nearest_phase_kernels(w: torch.Tensor, f: int):
    """nearest-x f upsample followed by a 3x3 'same' conv == f^2 convs of 2x2 taps on the source grid (f = 2: esrgan.upconv_phase_kernels;
    f = 3 here).  Output row 3 y + a reads upsampled rows 3 y + a - 1 .. 3 y + a + 1, i.e. source rows (y - 1, y, y) for a = 0, (y, y, y) for
    a = 1 and (y, y, y + 1) for a = 2: taps folded onto two source rows [r0, r1] with `pad` zero rows in front.  Returns
    [(phase index a * f + b, weight [cout][cin][2][2], (pad_top, pad_left))]."""
    if f == 2:
        return upconv_phase_kernels(w)
    assert f == 3
    zero = torch.zeros_like(w[:, :, 0])
    rows = {0: ([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 1), 1: ([w[:, :, 0] + w[:, :, 1] + w[:, :, 2], zero], 0), 2: ([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], 0)}
    out = []
    for a in range(3):
        (r0, r1), pad_t = rows[a]
        for b in range(3):
            def fold(r):  # r: [cout][cin][3] over kx
                if b == 0:
                    return torch.stack([r[:, :, 0], r[:, :, 1] + r[:, :, 2]], -1), 1
                if b == 1:
                    return torch.stack([r[:, :, 0] + r[:, :, 1] + r[:, :, 2], torch.zeros_like(r[:, :, 0])], -1), 0
                return torch.stack([r[:, :, 0] + r[:, :, 1], r[:, :, 2]], -1), 0
            (c0, pad_l), (c1, _) = fold(r0), fold(r1)
            out.append((a * 3 + b, torch.stack([c0, c1], 2).contiguous(), (pad_t, pad_l)))
    return out


def _head_specs(upsample: str, scale: int, dim: int, out_ch: int, mid: int, end_kernel: int) -> List[ParamSpec]:
    if scale == 1:
        return conv_specs('dim_to_in', dim, out_ch, 3)
    meta = torch.tensor([3, SAMPLE_MODS.index(upsample), scale, dim, out_ch, mid, 4], dtype=torch.uint8)
    specs: List[ParamSpec] = []
    if upsample == 'pixelshuffledirect':
        specs += conv_specs('dim_to_in.0', dim, out_ch * scale * scale, 3)
    elif upsample == 'pixelshuffle':
        specs += conv_specs('dim_to_in.0', dim, mid, 3)
        i = 2
        for r in ([3] if scale == 3 else [2] * int(math.log2(scale))):
            specs += conv_specs(f'dim_to_in.{i}', mid, r * r * mid, 3)
            i += 2
        specs += conv_specs(f'dim_to_in.{i}', mid, out_ch, 3)
    elif upsample == 'nearest+conv':  # [conv, Upsample(2), LeakyReLU] x n, conv, LeakyReLU, conv (arch.py:270-296); x3: one Upsample(3) step
        n = 1 if scale == 3 else int(math.log2(scale))
        for k in range(n + 1):
            specs += conv_specs(f'dim_to_in.{3 * k}', dim, dim, 3)
        specs += conv_specs(f'dim_to_in.{3 * n + 2}', dim, out_ch, 3)
    elif upsample == 'transpose+conv':  # ConvTranspose2d(4, 2, 1) [GELU ConvTranspose2d(4, 2, 1)] | ConvTranspose2d(3, 3, 0), then a 3x3 conv (arch.py:303-318)
        tw = lambda name, i, o, k: [(f'{name}.weight', (i, o, k, k), 'conv_w'), (f'{name}.bias', (o,), f'bias:{i * k * k}')]
        if scale == 4:
            specs += tw('dim_to_in.0', dim, dim, 4) + tw('dim_to_in.2', dim, out_ch, 4) + conv_specs('dim_to_in.3', out_ch, out_ch, 3)
        else:
            specs += tw('dim_to_in.0', dim, out_ch, 4 if scale == 2 else 3) + conv_specs('dim_to_in.1', out_ch, out_ch, 3)
    elif upsample == 'pa_up':  # [Upsample(2), conv, PA, LeakyReLU, conv, LeakyReLU] x n, conv (arch.py:325-352); x3: one Upsample(3) step
        n, cin = (1 if scale == 3 else int(math.log2(scale))), dim
        for k in range(n):
            specs += conv_specs(f'dim_to_in.{6 * k + 1}', cin, mid, 3) + conv_specs(f'dim_to_in.{6 * k + 2}.conv.0', mid, mid, 1) + conv_specs(f'dim_to_in.{6 * k + 4}', mid, mid, 3)
            cin = mid
        specs += conv_specs(f'dim_to_in.{6 * n}', mid, out_ch, 3)
        meta[3] = mid  # the reference records in_dim after its loop has re-bound it (arch.py:338)
    else:  # dysample
        i = 0
        if mid != dim:
            specs += conv_specs('dim_to_in.0', dim, mid, 3)
            i = 2
        if end_kernel != 1:
            raise NotImplementedError('GateRV3 DySample head: only a 1x1 end convolution is built (the sampling op fuses it)')
        specs += dysample_specs(f'dim_to_in.{i}', mid, out_ch, scale, 4)
    return specs + [('dim_to_in.MetaUpsample', meta, 'buffer_tensor')]


class GateRV3(EngineModule):
    def __init__(self, *, in_ch: int = 3, dim: int = 32, enc_blocks: Sequence[int] = (2, 2, 4, 8), dec_blocks: Sequence[int] = (2, 2, 2, 2),
                 num_latent: int = 12, scale: int = 1, upsample: str = 'pixelshuffle', upsample_mid_dim: int = 32, attention: bool = False,
                 span_blocks: int = 4, end_kernel: int = 1, seed: int = 0):
        if attention and (dim * 2 ** len(enc_blocks)) // ATT_HEADS > 32:
            raise NotImplementedError('GateRV3 latent attention: head_dim <= 32 (the channel-attention kernels)')
        if scale != 1 and upsample not in SUPPORTED_MODS:
            raise NotImplementedError(f'GateRV3 upsampler {upsample!r} is not built (supported: {SUPPORTED_MODS})')
        if scale != 1 and upsample == 'pixelshuffle' and scale & (scale - 1) and scale != 3:
            raise ValueError(f'scale {scale} is not supported. Supported scales: 2^n and 3.')
        if scale != 1 and upsample == 'transpose+conv' and scale not in (2, 3, 4):
            raise ValueError(f'scale {scale} is not supported. Supported scales: 2, 3, 4')
        if scale != 1 and upsample in ('nearest+conv', 'pa_up') and scale & (scale - 1) and scale != 3:
            raise ValueError(f'scale {scale} is not supported. Supported scales: 2^n and 3.')
        if len(enc_blocks) != len(dec_blocks) or dim % 16 or min(list(enc_blocks) + list(dec_blocks)) < 1:
            raise ValueError('GateRV3 needs as many decoder as encoder levels, >= 1 block per level and dim % 16 == 0 (planar-8 layout)')
        L = len(enc_blocks)
        specs: List[ParamSpec] = conv_specs('in_to_dim', in_ch, dim, 3)
        for i, nb in enumerate(enc_blocks):
            d = dim * 2 ** i
            for j in range(nb):
                specs += _meta_gated_specs(f'gater_encode.{i}.gated.{j}', d)
            specs += [(f'gater_encode.{i}.scale.0.weight', (d // 2, d, 3, 3), 'conv_w')]
        specs += _spab_specs('span_block0', dim)
        for k in range(span_blocks):
            specs += _spab_specs(f'span_n_b.{k}', dim)
        specs += _spab_specs('span_end', dim)
        specs += _conv3xc_specs('sisr_end_conv', dim, dim, 1, True) + conv_specs('sisr_cat_conv', 4 * dim, dim, 1)
        for k in range(num_latent):
            specs += _gated_cnn_specs(f'latent.{k}', dim * 2 ** L, attention)
        for i, nb in enumerate(dec_blocks):
            d = dim * 2 ** (L - i)
            specs += [(f'decode.{i}.scale.0.weight', (2 * d, d, 3, 3), 'conv_w')]
            for j in range(nb):
                specs += _meta_gated_specs(f'decode.{i}.gated.{j}', d // 2)
            specs += conv_specs(f'decode.{i}.shor', d, d // 2, 1)
        specs += [('gamma', (1, in_ch, 1, 1), 'affine_w')]
        specs += _head_specs(upsample, scale, dim, in_ch, upsample_mid_dim, end_kernel)
        super().__init__(specs, in_ch, in_ch, int(scale), seed=seed)
        self._plan_base_divisor = 2 ** L
        self.dim, self.enc_blocks, self.dec_blocks, self.num_latent = dim, list(enc_blocks), list(dec_blocks), num_latent
        self.scale, self.upsample, self.mid, self.span_blocks = int(scale), upsample, upsample_mid_dim, span_blocks
        self.attention = bool(attention)
        self.pad = 2 ** L

    # ------------------------------------------------------------------ host glue (arch.py:783-802)
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        xp = F.pad(x, (0, (self.pad - w % self.pad) % self.pad, 0, (self.pad - h % self.pad) % self.pad), 'reflect')
        y = super().forward(xp.contiguous())
        gamma = self.gamma.to(y.dtype)
        return y[:, :, : h * self.scale, : w * self.scale] + gamma * (F.interpolate(x, scale_factor=self.scale) if self.scale != 1 else x).to(y.dtype)

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        if h % self.pad == 0 and w % self.pad == 0:
            super().forward_into(x, out)
            out.add_(self.gamma.to(out.dtype) * (F.interpolate(x, scale_factor=self.scale) if self.scale != 1 else x).to(out.dtype))
            return out
        out.copy_(self.forward(x))
        return out

    # ------------------------------------------------------------------ plan
    def _gated_cnn(self, pb: PlanBuilder, w, p: str, x, out, s, d: int) -> None:
        """out = mish(fc2(mish(g) * cat(i, token_mix(c)))) with g, i, c = split(fc1(RMSNorm(x))) (arch.py:623-629)."""
        hidden = int(1.5 * d)
        # bf16 plan: the RMSNorm is folded into the 1x1 convs that consume it (statistics pass + rsb_conv_desc.ln_fold with a zero mean
        # term), the normalised map is never written
        if pb.compute_dtype == torch.bfloat16:
            pb.rmsnorm_stats(x, s['stats'], eps=1e-6)
            xn, fold = x, dict(ln=(s['stats'], w[f'{p}.norm.scale'], w[f'{p}.norm.offset']))
        else:
            pb.rmsnorm(x, s['xn'], w[f'{p}.norm.scale'], w[f'{p}.norm.offset'], eps=1e-6)
            xn, fold = s['xn'], {}
        w1, b1 = w[f'{p}.fc1.weight'], w[f'{p}.fc1.bias']
        g_rows, i_rows, c_rows = slice(0, hidden), slice(hidden, 2 * hidden - d), slice(2 * hidden - d, 2 * hidden)
        pb.conv(xn, s['ic'].slice(0, hidden - d), w1[i_rows], b1[i_rows], **fold)
        pb.conv(xn, s['c'], w1[c_rows], b1[c_rows], **fold)
        if f'{p}.token_mix.qkv.weight' in w:
            # Attention (arch.py:572-591): q, k normalised over the pixels, (q k^T) * temperature, softmax over channels, attn @ v —
            # the transposed attention of DAT's channel blocks (RSB_OP_CHANATTN) behind a 1x1 conv and a depthwise 3x3
            t = f'{p}.token_mix'
            pb.conv(s['c'], s['qkv'], w[f'{t}.qkv.weight'], None)
            pb.dwconv3(s['qkv'], s['qkvd'], w[f'{t}.qkv_dwconv.weight'], w[f'{t}.qkv_dwconv.bias'])
            pb.op(N.OP_CHANATTN, s['qkvd'], s['att'], d, ints=(ATT_HEADS, d), weights=(w[f'{t}.temperature'],))
            pb.conv(s['att'], s['ic'].slice(hidden - d, d), w[f'{t}.project_out.weight'], None)
        else:
            pb.dwconv(s['c'], s['ic'].slice(hidden - d, d), *merge_inception(w, f'{p}.token_mix', d))
        pb.conv(xn, s['gm'], w1[g_rows], b1[g_rows], act=N.ACT_MISH, combine=N.COMB_MUL, res1=s['ic'], **fold)
        pb.conv(s['gm'], out, w[f'{p}.fc2.weight'], w[f'{p}.fc2.bias'], act=N.ACT_MISH)

    def _meta_gated(self, pb: PlanBuilder, w, p: str, x, out, s, d: int) -> None:
        if pb.compute_dtype == torch.bfloat16:
            pb.rmsnorm_stats(x, s['stats'], eps=1e-6)
            pb.conv(x, s['h'], w[f'{p}.local.1.weight'], w[f'{p}.local.1.bias'], ln=(s['stats'], w[f'{p}.local.0.scale'], w[f'{p}.local.0.offset']))
        else:
            pb.rmsnorm(x, s['xn'], w[f'{p}.local.0.scale'], w[f'{p}.local.0.offset'], eps=1e-6)
            pb.conv(s['xn'], s['h'], w[f'{p}.local.1.weight'], w[f'{p}.local.1.bias'])
        # nn.Conv2d(2d, 2d, 3, groups=d): output channel o reads input channels 2 (o // 2), 2 (o // 2) + 1 -> block-diagonal dense
        # convs per chunk of <= 64 channels; SimpleGate multiplies the two halves of the OUTPUT: second half first
        wg, bg = w[f'{p}.local.2.weight'].double(), w[f'{p}.local.2.bias']
        chunk = next(c for c in (64, 48, 32, 16) if d % c == 0)
        for half in (1, 0):
            for c0 in range(0, d, chunk):
                o0 = half * d + c0
                dense = torch.zeros(chunk, chunk, 3, 3, dtype=torch.float64)
                for o in range(chunk):
                    dense[o, 2 * (o // 2):2 * (o // 2) + 2] = wg[o0 + o]
                kw = {} if half == 1 else dict(combine=N.COMB_MUL, res1=s['g2'].slice(c0, chunk))
                pb.conv(s['h'].slice(o0, chunk), (s['g2'] if half == 1 else s['gl']).slice(c0, chunk), dense, bg[o0:o0 + chunk], **kw)
        pb.chan_gate(s['gl'], s['xl'], x, w[f'{p}.sca.1.weight'], w[f'{p}.sca.1.bias'], w[f'{p}.gamma0'])   # x * sca(x) * gamma0 + short
        self._gated_cnn(pb, w, f'{p}.glob', s['xl'], s['m'], s, d)
        pb.chan_affine(s['m'], out, w[f'{p}.gamma1'], res=s['xl'])                                          # glob(x) * gamma1 + x

    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, L, full = self.dim, len(self.enc_blocks), self.pad
        grid = lambda l: full >> l
        cat4 = pb.buffer(4 * dim)  # [x | sisr | sisr_short | sisr_out] == the reference's torch.cat (arch.py:791)
        x0, sisr, sisr_short, sisr_out = (cat4.slice(i * dim, dim) for i in range(4))
        t2, p0, p1 = (pb.buffer(dim) for _ in range(3))
        pb.conv(INPUT, x0, w['in_to_dim.weight'], w['in_to_dim.bias'])

        def spab(prefix, src, dst, t1):
            pb.conv(src, t1, *merge_conv3xc_any(w, f'{prefix}.c1_r'), act=N.ACT_SILU)
            pb.conv(t1, t2, *merge_conv3xc_any(w, f'{prefix}.c2_r'), act=N.ACT_SILU)
            pb.conv(t2, dst, *merge_conv3xc_any(w, f'{prefix}.c3_r'), combine=N.COMB_SPAB_GATE, res1=src)

        t1 = pb.buffer(dim)
        spab('span_block0', x0, sisr_short, t1)
        cur = sisr_short
        for k in range(self.span_blocks):
            nxt = p0 if cur is not p0 else p1
            spab(f'span_n_b.{k}', cur, nxt, t1)
            cur = nxt
        nxt = p0 if cur is not p0 else p1
        spab('span_end', cur, nxt, sisr_out)  # second result = the (in-place activated) first conv output
        pb.conv(nxt, sisr, *merge_conv3xc_any(w, 'sisr_end_conv'))

        # per-level scratch (level l: width dim * 2^l on the H / 2^l grid), shared by the encoder and decoder blocks of the level
        def scratch(d, g, meta=True):
            hidden = int(1.5 * d)
            s = dict(xn=pb.buffer(d, scale=g), stats=pb.buffer(8, scale=g), ic=pb.buffer(hidden, scale=g), c=pb.buffer(d, scale=g), gm=pb.buffer(hidden, scale=g),
                     a=pb.buffer(d, scale=g), b=pb.buffer(d, scale=g))
            if meta:
                s.update(h=pb.buffer(2 * d, scale=g), g2=pb.buffer(d, scale=g), gl=pb.buffer(d, scale=g), xl=pb.buffer(d, scale=g), m=pb.buffer(d, scale=g))
            return s

        levels = [scratch(dim * 2 ** l, grid(l)) for l in range(L)]
        cats = [pb.buffer(2 * dim * 2 ** l, scale=grid(l)) for l in range(L)]  # decoder concat: [upsampled | encoder output of the level]
        cur = x0
        for l, nb in enumerate(self.enc_blocks):
            d, s = dim * 2 ** l, levels[l]
            for j in range(nb):
                dst = cats[l].slice(d, d) if j == nb - 1 else (s['a'] if cur is not s['a'] else s['b'])
                self._meta_gated(pb, w, f'gater_encode.{l}.gated.{j}', cur, dst, s, d)
                cur = dst
            # Down: conv d -> d / 2 (no bias) + PixelUnshuffle(2) -> 2 d channels on the next grid
            half = pb.buffer(d // 2, scale=grid(l))
            pb.conv(cur, half, w[f'gater_encode.{l}.scale.0.weight'], None)
            u5 = pb.buffer(5 * (d // 2), scale=grid(l + 1))
            pb.unshuffle_pool(half, u5)
            cur = u5.slice(0, 2 * d)
        dl = dim * 2 ** L
        s = scratch(dl, grid(L), meta=False)
        if self.attention:
            s.update(qkv=pb.buffer(3 * dl, scale=grid(L)), qkvd=pb.buffer(3 * dl, scale=grid(L)), att=pb.buffer(dl, scale=grid(L)))
        for k in range(self.num_latent):
            dst = s['a'] if cur is not s['a'] else s['b']
            self._gated_cnn(pb, w, f'latent.{k}', cur, dst, s, dl)
            cur = dst
        for i, nb in enumerate(self.dec_blocks):
            l = L - 1 - i
            d, s = dim * 2 ** l, levels[l]  # this level's width; the block's input has 2 d channels on the coarser grid
            up = pb.buffer(4 * d, scale=grid(l + 1))
            pb.conv(cur, up, w[f'decode.{i}.scale.0.weight'], None)
            pb.se_shuffle(up, cats[l].slice(0, d))  # PixelShuffle(2)
            cur = s['a']
            pb.conv(cats[l], cur, w[f'decode.{i}.shor.weight'], w[f'decode.{i}.shor.bias'])
            for j in range(nb):
                dst = s['a'] if cur is not s['a'] else s['b']
                self._meta_gated(pb, w, f'decode.{i}.gated.{j}', cur, dst, s, d)
                cur = dst
        # x + sisr, sisr = sisr_cat_conv(cat): the 1x1 conv runs last with the decoder output as its residual
        xs = levels[0]['xn']
        pb.conv(cat4, xs, w['sisr_cat_conv.weight'], w['sisr_cat_conv.bias'], combine=N.COMB_AXPY, res1=cur)
        self._head(pb, w, xs)

    def _head(self, pb: PlanBuilder, w, x) -> None:
        r, full = self.scale, self.pad
        if r == 1:
            pb.conv(x, OUTPUT, w['dim_to_in.weight'], w['dim_to_in.bias'], ps=1)
        elif self.upsample == 'pixelshuffledirect':
            pb.conv(x, OUTPUT, w['dim_to_in.0.weight'], w['dim_to_in.0.bias'], ps=r)
        elif self.upsample == 'pixelshuffle':
            mid = self.mid
            cur = pb.buffer(mid)
            pb.conv(x, cur, w['dim_to_in.0.weight'], w['dim_to_in.0.bias'], act=N.ACT_LRELU, act_param=0.01)
            grid, i = 1, 2
            for f in ([3] if r == 3 else [2] * int(math.log2(r))):
                nxt = pb.buffer(mid, scale=full * grid * f)
                wk, bk = w[f'dim_to_in.{i}.weight'], w[f'dim_to_in.{i}.bias']
                perm = torch.arange(mid * f * f).view(mid, f * f).t().reshape(-1)  # PixelShuffle: channel c * f^2 + phase
                for phase in range(f * f):
                    sel = perm[phase * mid:(phase + 1) * mid]
                    pb.conv(cur, nxt, wk[sel], bk[sel], dst_ps=f, dst_phase=phase)
                cur, grid, i = nxt, grid * f, i + 2
            pb.conv(cur, OUTPUT, w[f'dim_to_in.{i}.weight'], w[f'dim_to_in.{i}.bias'], ps=1)
        elif self.upsample == 'nearest+conv':
            # lrelu commutes with the nearest upsample: z0 = lrelu(conv0(x)) on the input grid, every later conv reads an upsampled map
            # = four 2x2 phase convs on the source grid (esrgan.upconv_phase_kernels), stored pixel-interleaved
            f = 3 if r == 3 else 2
            n, lrelu = (1 if r == 3 else int(math.log2(r))), dict(act=N.ACT_LRELU, act_param=0.2)
            cur = pb.buffer(self.dim)
            pb.conv(x, cur, w['dim_to_in.0.weight'], w['dim_to_in.0.bias'], **lrelu)
            grid = 1
            for k in range(1, n + 1):
                nxt = pb.buffer(self.dim, scale=full * grid * f)
                for phase, wk, pad2 in nearest_phase_kernels(w[f'dim_to_in.{3 * k}.weight'], f):
                    pb.conv(cur, nxt, wk, w[f'dim_to_in.{3 * k}.bias'], dst_ps=f, dst_phase=phase, pad=pad2, **lrelu)
                cur, grid = nxt, grid * f
            pb.conv(cur, OUTPUT, w[f'dim_to_in.{3 * n + 2}.weight'], w[f'dim_to_in.{3 * n + 2}.bias'], ps=1)
        elif self.upsample == 'transpose+conv':
            oc = self.out_channels

            def transposed(src, dst, name, f, **kw):
                """ConvTranspose2d as sub-pixel phase convs on the source grid.  k = 4, s = 2, p = 1: output row 2 y + a gets input rows
                (y - 1, y) through kernel rows (3, 1) for a = 0 and rows (y, y + 1) through kernel rows (2, 0) for a = 1 (columns alike)
                -> four 2x2 convs; k = 3, s = 3, p = 0: every output pixel has one source pixel -> nine 1x1 convs."""
                wt, bt = w[f'{name}.weight'], w[f'{name}.bias']  # [in][out][k][k]
                co = -(-wt.shape[1] // 8) * 8  # a sub-pixel phase writes whole 8-channel planes: zero output channels fill the last one
                wt, bt = F.pad(wt, (0, 0, 0, 0, 0, co - wt.shape[1])), F.pad(bt, (0, co - bt.shape[0]))
                dst = dst.slice(0, co) if dst.channels >= co else dst
                if f == 3:
                    for phase in range(9):
                        pb.conv(src, dst, wt[:, :, phase // 3, phase % 3].t().reshape(wt.shape[1], wt.shape[0], 1, 1), bt, dst_ps=3, dst_phase=phase, **kw)
                    return
                taps = {0: ((3, 1), 1), 1: ((2, 0), 0)}  # parity -> (kernel indices of the two taps, zero padding in front)
                for a in (0, 1):
                    for b in (0, 1):
                        (ky, pt), (kx, pl) = taps[a], taps[b]
                        wk = wt[:, :, list(ky)][:, :, :, list(kx)].permute(1, 0, 2, 3).contiguous()
                        pb.conv(src, dst, wk, bt, dst_ps=2, dst_phase=a * 2 + b, pad=(pt, pl), **kw)

            hi = pb.buffer(16, scale=full * r)  # the head's out_ch-channel map on the output grid (16 channels: two whole planes for the conv)
            if r == 4:
                mid = pb.buffer(self.dim, scale=full * 2)
                transposed(x, mid, 'dim_to_in.0', 2, act=N.ACT_GELU)
                transposed(mid, hi, 'dim_to_in.2', 2)
                last = 'dim_to_in.3'
            else:
                transposed(x, hi, 'dim_to_in.0', r)
                last = 'dim_to_in.1'
            wl = torch.zeros(oc, 16, 3, 3, dtype=w[f'{last}.weight'].dtype)
            wl[:, :oc] = w[f'{last}.weight']
            pb.conv(hi, OUTPUT, wl, w[f'{last}.bias'], ps=1)
        elif self.upsample == 'pa_up':
            f = 3 if r == 3 else 2
            n, mid, cur, grid = (1 if r == 3 else int(math.log2(r))), self.mid, x, 1
            for k in range(n):
                g2 = full * grid * f
                a, t, u, v = (pb.buffer(mid, scale=g2) for _ in range(4))
                for phase, wk, pad2 in nearest_phase_kernels(w[f'dim_to_in.{6 * k + 1}.weight'], f):
                    pb.conv(cur, a, wk, w[f'dim_to_in.{6 * k + 1}.bias'], dst_ps=f, dst_phase=phase, pad=pad2)
                # PA: a * sigmoid(conv1x1(a)) as the 1x1 conv's epilogue; the LeakyReLU behind the product is one more 1x1 pass
                pb.conv(a, t, w[f'dim_to_in.{6 * k + 2}.conv.0.weight'], w[f'dim_to_in.{6 * k + 2}.conv.0.bias'], act=N.ACT_SIGMOID, combine=N.COMB_MUL, res1=a)
                pb.conv(t, u, torch.eye(mid).view(mid, mid, 1, 1), None, act=N.ACT_LRELU, act_param=0.2)
                pb.conv(u, v, w[f'dim_to_in.{6 * k + 4}.weight'], w[f'dim_to_in.{6 * k + 4}.bias'], act=N.ACT_LRELU, act_param=0.2)
                cur, grid = v, grid * f
            pb.conv(cur, OUTPUT, w[f'dim_to_in.{6 * n}.weight'], w[f'dim_to_in.{6 * n}.bias'], ps=1)
        else:  # dysample
            i = 0
            if self.mid != self.dim:
                cur = pb.buffer(self.mid)
                pb.conv(x, cur, w['dim_to_in.0.weight'], w['dim_to_in.0.bias'], act=N.ACT_LRELU, act_param=0.01)
                x, i = cur, 2
            emit_dysample(pb, w, f'dim_to_in.{i}', x, self.out_channels, r, 4)


class GateRV3Arch(Architecture[GateRV3]):
    def __init__(self):
        mg = ('gamma0', 'gamma1', 'local.0.scale', 'local.0.offset', 'local.1.weight', 'local.1.bias', 'local.2.weight', 'local.2.bias',
              'sca.1.weight', 'sca.1.bias', 'glob.norm.scale', 'glob.norm.offset', 'glob.fc1.weight', 'glob.fc1.bias',
              'glob.token_mix.dwconv_hw.weight', 'glob.token_mix.dwconv_hw.bias', 'glob.token_mix.dwconv_w.weight', 'glob.token_mix.dwconv_w.bias',
              'glob.token_mix.dwconv_h.weight', 'glob.token_mix.dwconv_h.bias', 'glob.fc2.weight', 'glob.fc2.bias')
        c3 = ('sk.weight', 'conv.0.weight', 'conv.1.weight', 'conv.2.weight', 'eval_conv.weight')
        keys = ['in_to_dim.weight', 'in_to_dim.bias'] + [f'gater_encode.0.gated.0.{k}' for k in mg] + ['gater_encode.0.scale.0.weight']
        keys += [f'{b}.{c}.{k}' for b in ('span_block0', 'span_n_b.0', 'span_end') for c in ('c1_r', 'c2_r', 'c3_r') for k in c3]
        keys += [f'sisr_end_conv.{k}' for k in ('sk.weight', 'sk.bias', 'conv.0.weight', 'conv.0.bias', 'conv.1.weight', 'conv.1.bias', 'conv.2.weight',
                                                'conv.2.bias', 'eval_conv.weight', 'eval_conv.bias')]
        keys += ['sisr_cat_conv.weight', 'sisr_cat_conv.bias', 'decode.0.scale.0.weight'] + [f'decode.0.gated.0.{k}' for k in mg]
        keys += ['decode.0.shor.weight', 'decode.0.shor.bias']
        super().__init__(uid='GateRV3', detect=KeyCondition.has_all(*keys))

    def load(self, state: Mapping[str, object]):
        # gaterv3/__init__.py:122-157
        dim, in_ch = state['in_to_dim.weight'].shape[:2]
        enc_blocks = [get_seq_len(state, f'gater_encode.{i}.gated') for i in range(get_seq_len(state, 'gater_encode'))]
        latent = get_seq_len(state, 'latent')
        dec_blocks = [get_seq_len(state, f'decode.{i}.gated') for i in range(get_seq_len(state, 'decode'))]
        end_kernel = 1
        if 'dim_to_in.MetaUpsample' in state:
            _, index, scale, _, out_ch, upsample_dim, _ = [int(v) for v in state['dim_to_in.MetaUpsample']]
            upsampler = SAMPLE_MODS[index]
            if upsampler == 'dysample' and 'dim_to_in.0.weight' not in state:
                upsample_dim = dim
                end_kernel = state['dim_to_in.0.end_conv.weight'].shape[2]
            elif upsampler == 'dysample':
                end_kernel = state['dim_to_in.2.end_conv.weight'].shape[2]
        else:
            scale, upsample_dim, upsampler = 1, 32, 'conv'
        attention = 'latent.0.token_mix.qkv_dwconv.weight' in state
        model = GateRV3(in_ch=in_ch, dim=dim, enc_blocks=enc_blocks, dec_blocks=dec_blocks, num_latent=latent, scale=scale, upsample=upsampler,
                        upsample_mid_dim=upsample_dim, attention=attention, span_blocks=get_seq_len(state, 'span_n_b'), end_kernel=end_kernel)
        return self._enhance_model(model=model, in_channels=in_ch, out_channels=int(in_ch), upscale=scale, name='GateRV3')
