"""SwinIR on the B200 engine — all four reconstruction heads, '1conv' and '3conv' residual connections.

Reference: /root/reference/resselt/archs/swinir/arch.py:735-1011 (model), :192-335 (SwinTransformerBlock), :75-170
(WindowAttention), :17-40 (Mlp), :499-593 (RSTB), :689-732 (Upsample / UpsampleOneStep), loader
/root/reference/resselt/archs/swinir/__init__.py:9-119, padding /root/reference/resselt/utilities/padding.py:5-29.

Lowering (token == pixel of a planar-8 buffer):
  * every nn.Linear / nn.Conv2d is a tensor-core conv op (1x1 for the linears); ``shortcut + attn``, ``x + mlp(...)``,
    ``conv(...) + x`` of the RSTB and ``conv_after_body(...) + feat`` are their epilogues, GELU after fc1 too;
  * W-MSA / SW-MSA is the engine's fused window-attention kernel with square windows: the learned
    ``relative_position_bias_table`` ([(2w-1)^2][heads], gathered through ``relative_position_index`` in the reference,
    arch.py:150-158) is already the per-offset table the kernel indexes by (dy + w - 1)(2w - 1) + (dx + w - 1);
    cyclic shift, window partition / reverse and the shift mask (arch.py:268-293, 305-332) are addressing inside it;
  * a residual group never copies its input: block 0 reads buffer A and writes B, the other blocks update B in place, the
    group's conv writes C = conv(B) + A, and the three buffers rotate;
  * ``check_image_size`` (reflect padding to a multiple of the window, arch.py:944-945) and the final crop (:1011) are
    host-side glue around the plan, like the reference does them around its own forward.
"""
from __future__ import annotations

import math
from typing import Mapping

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_pixelshuffle_params, get_seq_len
from ._common import HEAD_PAD, conv_specs, emit_resi_conv, pad_head_cols, pad_head_rows, resi_conv_specs, winattn_head_padded
from .dat import RGB_MEAN, _lin_specs, _ln_specs
from .esrgan import upconv_phase_kernels

UPSAMPLERS = ('pixelshuffle', 'pixelshuffledirect', 'nearest+conv', '')
NUM_FEAT = 64  # arch.py:795


def _relative_position_index(ws: int) -> torch.Tensor:
    """WindowAttention.__init__ (arch.py:117-127)."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _shift_mask(size: int, ws: int, shift: int) -> torch.Tensor:
    """attn_mask buffer of a shifted block (arch.py:268-293); only stored, never read by the engine."""
    img = torch.zeros(size, size)
    cnt = 0
    for a in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for b in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[a, b] = cnt
            cnt += 1
    win = img.view(size // ws, ws, size // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class SwinIR(EngineModule):
    def __init__(
        self,
        img_size: int = 64,
        in_chans: int = 3,
        embed_dim: int = 180,
        depths=(6, 6, 6, 6, 6, 6),
        num_heads=(6, 6, 6, 6, 6, 6),
        window_size: int = 8,
        mlp_ratio: float = 2.0,
        qkv_bias: bool = True,
        upscale: int = 2,
        img_range: float = 1.0,
        upsampler: str = 'pixelshuffle',
        resi_connection: str = '1conv',
        seed: int = 0,
    ):
        if upsampler not in UPSAMPLERS or resi_connection not in ('1conv', '3conv'):
            raise ValueError(f'unknown upsampler {upsampler!r} / resi_connection {resi_connection!r}')
        if img_size <= window_size or img_size % window_size:
            # the reference then shrinks the window and drops the shift (arch.py:233-236); no released model does this
            raise NotImplementedError('SwinIR needs img_size to be a multiple of (and larger than) window_size')
        dim, hidden, ws = embed_dim, int(embed_dim * mlp_ratio), window_size
        if upsampler == 'pixelshuffle' and upscale & (upscale - 1) and upscale != 3:
            raise ValueError(f'scale {upscale} is not supported. Supported scales: 2^n and 3.')  # arch.py:707
        if upsampler == 'nearest+conv' and upscale not in (2, 4, 8):
            raise NotImplementedError('nearest+conv head: x2, x4 or x8')
        if upsampler == '':
            upscale = 1
        specs = conv_specs('conv_first', in_chans, dim, 3) + _ln_specs('patch_embed.norm', dim)
        for i, (nblk, heads) in enumerate(zip(depths, num_heads)):
            if dim % heads or heads % 2 or dim // heads > 32:
                raise NotImplementedError('window attention kernel: even head count, head_dim <= 32')
            for b in range(nblk):
                p = f'layers.{i}.residual_group.blocks.{b}'
                if b % 2 == 1:
                    specs += [(f'{p}.attn_mask', _shift_mask(img_size, ws, ws // 2), 'buffer_tensor')]
                specs += _ln_specs(f'{p}.norm1', dim)
                specs += [(f'{p}.attn.relative_position_bias_table', ((2 * ws - 1) ** 2, heads), 'normal:0.3'),
                          (f'{p}.attn.relative_position_index', _relative_position_index(ws), 'buffer_tensor')]
                specs += _lin_specs(f'{p}.attn.qkv', dim, 3 * dim, qkv_bias) + _lin_specs(f'{p}.attn.proj', dim, dim)
                specs += _ln_specs(f'{p}.norm2', dim)
                specs += _lin_specs(f'{p}.mlp.fc1', dim, hidden) + _lin_specs(f'{p}.mlp.fc2', hidden, dim)
            specs += resi_conv_specs(f'layers.{i}.conv', dim, resi_connection)
        specs += _ln_specs('norm', dim) + resi_conv_specs('conv_after_body', dim, resi_connection)
        if upsampler == 'pixelshuffle':
            specs += conv_specs('conv_before_upsample.0', dim, NUM_FEAT, 3)
            steps = [3] if upscale == 3 else [2] * int(math.log2(upscale))
            for i, r in enumerate(steps):
                specs += conv_specs(f'upsample.{2 * i}', NUM_FEAT, r * r * NUM_FEAT, 3)
            specs += conv_specs('conv_last', NUM_FEAT, in_chans, 3)
        elif upsampler == 'pixelshuffledirect':
            specs += conv_specs('upsample.0', dim, upscale * upscale * in_chans, 3)
        elif upsampler == 'nearest+conv':
            specs += conv_specs('conv_before_upsample.0', dim, NUM_FEAT, 3)
            for n in range(1, int(math.log2(upscale)) + 1):
                specs += conv_specs(f'conv_up{n}', NUM_FEAT, NUM_FEAT, 3)
            specs += conv_specs('conv_hr', NUM_FEAT, NUM_FEAT, 3, gain=2.0) + conv_specs('conv_last', NUM_FEAT, in_chans, 3, gain=3.0)
        else:
            specs += conv_specs('conv_last', dim, in_chans, 3)
        super().__init__(specs, in_chans, in_chans, upscale, seed=seed)
        self.dim, self.hidden, self.window_size = dim, hidden, ws
        self.depths, self.heads = list(depths), list(num_heads)
        self.img_range, self.img_size = float(img_range), img_size
        self.upsampler, self.resi_connection = upsampler, resi_connection

    # ------------------------------------------------------------------ plan
    def _resi_conv(self, pb: PlanBuilder, w, name: str, src, dst, res, tmp_a, tmp_b) -> None:
        emit_resi_conv(pb, w, name, self.resi_connection, src, dst, res, tmp_a, tmp_b)

    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, hidden, ws = self.dim, self.hidden, self.window_size
        lin_w = lambda name: w[f'{name}.weight'].view(*w[f'{name}.weight'].shape, 1, 1)
        lin_b = lambda name: w.get(f'{name}.bias')
        # bf16 plan, 8x8 windows, head_dim < 32: heads on 32-channel strides -> the tcgen05 window-attention kernel (winattn_tc.cu);
        # the padding is zero rows of the qkv weights and zero columns of proj, so it costs no pass
        padded = {h: winattn_head_padded(pb.compute_dtype, dim, h, (ws, ws)) for h in set(self.heads)}
        apad = max([dim] + [h * HEAD_PAD for h in padded if padded[h]])
        feat, xn, att = pb.buffer(dim), pb.buffer(dim), pb.buffer(apad)
        a, b, c = pb.buffer(dim), pb.buffer(dim), pb.buffer(dim)
        pad = max((dim + 15) // 16 * 16, apad)  # q | k | v start on 16-channel boundaries
        # wide MLPs: hidden width padded with zero weights to a multiple of 64 so fc2 stages whole 64-channel K chunks
        hpad = hidden if hidden <= 128 else (hidden + 63) // 64 * 64
        qkv, hid = pb.buffer(3 * pad), pb.buffer(hpad)
        stats = pb.buffer(8)  # LayerNorm statistics of norm1 / norm2, folded into the linears that consume them
        tmp_a = tmp_b = None
        if self.resi_connection == '3conv':
            tmp_a, tmp_b = pb.buffer(dim // 4), pb.buffer(dim // 4)
        mean = RGB_MEAN if self.in_channels == 3 else (0.0, 0.0, 0.0)
        pb.conv(INPUT, feat, w['conv_first.weight'], w['conv_first.bias'], in_mean=mean, in_scale=self.img_range)
        pb.layernorm(feat, a, w['patch_embed.norm.weight'], w['patch_embed.norm.bias'])
        for i, (nblk, heads) in enumerate(zip(self.depths, self.heads)):
            cur = a
            # bf16 plan: the residual linears (`shortcut + proj(..)`, `x + fc2(..)`) also write the per-pixel {sum, sum of squares} of what
            # they store (rsb_conv_desc.ln_out), so norm2 and the next block's norm1 need no statistics pass of their own
            w2_bytes = ((hpad + 15) // 16 * 16) * ((dim + 15) // 16 * 16) * 2
            fuse_stats = pb.ln_out_supported(dim) and w2_bytes <= 150 * 1024
            raw = False  # does `stats` hold raw sums (written by a conv) or the statistics op's {rstd, -mean * rstd}?
            for blk in range(nblk):
                p = f'layers.{i}.residual_group.blocks.{blk}'
                if not raw:
                    pb.layernorm_stats(cur, stats)  # norm1 -> qkv: LayerNorm applied in the linears' epilogues, its output never written
                ln1 = (stats, w[f'{p}.norm1.weight'], w[f'{p}.norm1.bias']) + ((1e-5,) if raw else ())
                wq, bq = lin_w(f'{p}.attn.qkv'), lin_b(f'{p}.attn.qkv')
                hp = padded[heads]
                width = heads * HEAD_PAD if hp else dim
                for part in range(3):  # one conv per q / k / v (UMMA N <= 256)
                    rows = slice(part * dim, (part + 1) * dim)
                    wp, bp = wq[rows], None if bq is None else bq[rows]
                    if hp:
                        wp, bp = pad_head_rows(wp, dim, heads), pad_head_rows(bp, dim, heads)
                    pb.conv(cur, qkv.slice(part * pad, width), wp, bp, ln=ln1)
                table = w[f'{p}.attn.relative_position_bias_table']  # [(2w-1)^2][heads]; the kernel wants one table per head half
                pb.op(N.OP_WINATTN, qkv, att, dim, ints=(heads, ws, ws, blk % 2, pad, HEAD_PAD if hp else 0), floats=((dim // heads) ** -0.5,),
                      weights=(table[:, : heads // 2].contiguous(), table[:, heads // 2:].contiguous()))
                wproj = lin_w(f'{p}.attn.proj')
                pb.conv(att.slice(0, width), b, pad_head_cols(wproj, dim, heads) if hp else wproj, lin_b(f'{p}.attn.proj'),
                        combine=N.COMB_AXPY, res1=cur, ln_out=stats if fuse_stats else None)  # shortcut + attn
                cur = b
                if not fuse_stats:
                    pb.layernorm_stats(cur, stats)  # norm2 -> fc1
                ln2 = (stats, w[f'{p}.norm2.weight'], w[f'{p}.norm2.bias']) + ((1e-5,) if fuse_stats else ())
                w1, b1, w2 = lin_w(f'{p}.mlp.fc1'), lin_b(f'{p}.mlp.fc1'), lin_w(f'{p}.mlp.fc2')
                if hpad != hidden:  # gelu(0 * x + 0) = 0 feeds zero fc2 columns
                    w1, b1 = F.pad(w1, (0, 0, 0, 0, 0, 0, 0, hpad - hidden)), F.pad(b1, (0, hpad - hidden))
                    w2 = F.pad(w2, (0, 0, 0, 0, 0, hpad - hidden))
                pb.conv(cur, hid, w1, b1, act=N.ACT_GELU, ln=ln2)
                raw = fuse_stats and blk + 1 < nblk  # the next block's norm1 statistics
                pb.conv(hid, cur, w2, lin_b(f'{p}.mlp.fc2'), combine=N.COMB_AXPY, res1=cur, ln_out=stats if raw else None)      # x + mlp(norm2(x))
            self._resi_conv(pb, w, f'layers.{i}.conv', cur, c, a, tmp_a, tmp_b)  # RSTB: conv(blocks(x)) + x
            a, c = c, a
        pb.layernorm(a, xn, w['norm.weight'], w['norm.bias'])
        att = att.slice(0, dim)
        self._resi_conv(pb, w, 'conv_after_body', xn, att, feat, tmp_a, tmp_b)
        out_kw = dict(out_scale=1.0 / self.img_range, out_mean=mean)
        if self.upsampler == 'pixelshuffledirect':
            pb.conv(att, OUTPUT, w['upsample.0.weight'], w['upsample.0.bias'], ps=self.upscale, **out_kw)
            return
        if self.upsampler == '':
            # x + conv_last(res), then / img_range + mean (arch.py:1004-1009)  ==  raw input + conv_last(res) / img_range
            pb.conv(att, OUTPUT, w['conv_last.weight'] / self.img_range, w['conv_last.bias'] / self.img_range, ps=1, add_base=True)
            return
        cur = pb.buffer(NUM_FEAT)
        pb.conv(att, cur, w['conv_before_upsample.0.weight'], w['conv_before_upsample.0.bias'], act=N.ACT_LRELU, act_param=0.01)
        grid = 1
        if self.upsampler == 'pixelshuffle':
            steps = [3] if self.upscale == 3 else [2] * int(math.log2(self.upscale))
            for i, r in enumerate(steps):
                nxt = pb.buffer(NUM_FEAT, scale=grid * r)
                # conv 64 -> 64 r^2 + PixelShuffle(r): channel c * r^2 + phase -> one conv per sub-pixel phase
                wk, bk = w[f'upsample.{2 * i}.weight'], w[f'upsample.{2 * i}.bias']
                perm = torch.arange(NUM_FEAT * r * r).view(NUM_FEAT, r * r).t().reshape(-1)
                for phase in range(r * r):
                    sel = perm[phase * NUM_FEAT:(phase + 1) * NUM_FEAT]
                    pb.conv(cur, nxt, wk[sel], bk[sel], dst_ps=r, dst_phase=phase)
                cur, grid = nxt, grid * r
        else:  # nearest+conv: lrelu(conv_up(nearest x2)) as four 2x2 phase convs on the low-res grid
            lrelu = dict(act=N.ACT_LRELU, act_param=0.2)
            for n in range(1, int(math.log2(self.upscale)) + 1):
                nxt = pb.buffer(NUM_FEAT, scale=grid * 2)
                for phase, wk, pad2 in upconv_phase_kernels(w[f'conv_up{n}.weight']):
                    pb.conv(cur, nxt, wk, w[f'conv_up{n}.bias'], dst_ps=2, dst_phase=phase, pad=pad2, **lrelu)
                cur, grid = nxt, grid * 2
            hr = pb.buffer(NUM_FEAT, scale=grid)
            pb.conv(cur, hr, w['conv_hr.weight'], w['conv_hr.bias'], **lrelu)
            cur = hr
        pb.conv(cur, OUTPUT, w['conv_last.weight'], w['conv_last.bias'], ps=1, **out_kw)

    # ------------------------------------------------------------------ check_image_size + crop (host-side glue)
    def _padded(self, x: torch.Tensor) -> torch.Tensor:
        ws = self.window_size
        h, w = x.shape[-2:]
        if h % ws == 0 and w % ws == 0:
            return x
        return F.pad(x, (0, (ws - w % ws) % ws, 0, (ws - h % ws) % ws), 'reflect')

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        xp = self._padded(x)
        y = super().forward(xp)
        return y if xp is x else y[:, :, : h * self.upscale, : w * self.upscale]

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if self._padded(x) is x:
            return super().forward_into(x, out)
        out.copy_(self.forward(x))
        return out


class SwinIRArch(Architecture[SwinIR]):
    def __init__(self):
        super().__init__(
            uid='SwinIR',
            detect=KeyCondition.has_all(
                'layers.0.residual_group.blocks.0.norm1.weight',
                'conv_first.weight',
                'layers.0.residual_group.blocks.0.mlp.fc1.bias',
                'layers.0.residual_group.blocks.0.attn.relative_position_index',
            ),
        )

    def load(self, state_dict: Mapping[str, object]):
        if 'conv_before_upsample.0.weight' in state_dict:
            upsampler = 'nearest+conv' if 'conv_up1.weight' in state_dict else 'pixelshuffle'
        elif 'upsample.0.weight' in state_dict:
            upsampler = 'pixelshuffledirect'
        else:
            upsampler = ''
        if 'conv_first.1.weight' in state_dict:
            # bicubic-upsample + pixel-unshuffle front end (swinir/__init__.py:45-48, arch.py:969-972)
            raise NotImplementedError('SwinIR checkpoints with a pixel-unshuffle stem (conv_first.1.*) are not supported')
        in_ch = state_dict['conv_first.weight'].shape[1]
        out_ch = state_dict['conv_last.weight'].shape[0] if 'conv_last.weight' in state_dict else in_ch
        if out_ch != in_ch:
            raise NotImplementedError('SwinIR with num_out_ch != num_in_ch')  # the reference model cannot express it either (arch.py:794)
        upscale = 1
        if upsampler == 'nearest+conv':
            upscale = 2 ** sum(1 for k in state_dict if 'conv_up' in k and 'bias' not in k)
        elif upsampler == 'pixelshuffle':
            upscale, _ = get_pixelshuffle_params(state_dict, 'upsample')
        elif upsampler == 'pixelshuffledirect':
            upscale = int(math.sqrt(state_dict['upsample.0.bias'].shape[0] // out_ch))
        embed_dim = state_dict['conv_first.weight'].shape[0]
        blk0 = 'layers.0.residual_group.blocks.0'
        mlp_ratio = float(state_dict[f'{blk0}.mlp.fc1.bias'].shape[0] / embed_dim)
        window_size = int(math.sqrt(state_dict[f'{blk0}.attn.relative_position_index'].shape[0]))
        img_size = 64
        if 'layers.0.residual_group.blocks.1.attn_mask' in state_dict:
            img_size = int(math.sqrt(state_dict['layers.0.residual_group.blocks.1.attn_mask'].shape[0]) * window_size)
        num_layers = get_seq_len(state_dict, 'layers')
        depths = [get_seq_len(state_dict, f'layers.{i}.residual_group.blocks') for i in range(num_layers)]
        num_heads = [state_dict[f'layers.{i}.residual_group.blocks.0.attn.relative_position_bias_table'].shape[1] for i in range(num_layers)]
        resi_connection = '1conv' if 'conv_after_body.weight' in state_dict else '3conv'
        img_range = 255.0 if window_size == 7 else 1.0  # the JPEG models (swinir/__init__.py:91-92)
        model = SwinIR(
            img_size=img_size, in_chans=in_ch, embed_dim=embed_dim, depths=depths, num_heads=num_heads, window_size=window_size,
            mlp_ratio=mlp_ratio, qkv_bias=f'{blk0}.attn.qkv.bias' in state_dict, upscale=upscale, img_range=img_range,
            upsampler=upsampler, resi_connection=resi_connection,
        )
        return self._enhance_model(model=model, in_channels=in_ch, out_channels=out_ch, upscale=upscale, name='SwinIR')
