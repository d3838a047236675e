"""DAT (Dual Aggregation Transformer) on the B200 engine — 'pixelshuffle' and 'pixelshuffledirect' heads, '1conv' / '3conv'.

Reference: /root/reference/resselt/archs/dat/arch.py:828-990 (model), :615-683 (DATB), :270-513 (Adaptive_Spatial_Attention),
:516-612 (Adaptive_Channel_Attention), :40-101 (SGFN / SpatialGate), :104-143 (DynamicPosBias), :686-780 (ResidualGroup),
loader /root/reference/resselt/archs/dat/__init__.py:9-105.

Lowering (token == pixel of a planar-8 buffer, channels padded 180 -> 192 with zero weights):
  * every nn.Linear / nn.Conv2d is a tensor-core conv op (1x1 for the linears); residual adds are their epilogues
    (``x + proj(...)``, ``x + fc2(...)``, ``res + conv(...)``, ``conv_after_body(...) + feat``), GELU after fc1 too;
  * the dynamic position-bias MLP depends only on weights: it is evaluated once per plan on the host into a
    [(2Hs-1)(2Ws-1)][heads/2] table per branch; roll / pad / window partition / mask are addressing inside the fused
    window-attention kernel (the reference rebuilds the mask on the CPU each forward when H != img_size, arch.py:471-474);
  * eval-mode BatchNorm is folded into the preceding conv on the host.
"""
from __future__ import annotations

import math
from typing import List, Mapping

import torch
import torch.nn.functional as F

from ..engine import INPUT, OUTPUT, EngineModule, PlanBuilder
from ..engine import native as N
from ..factory import Architecture, KeyCondition
from ..utilities.state_dict import get_seq_len, pixelshuffle_scale
from ._common import HEAD_PAD, conv_specs, emit_resi_conv, pad_head_cols, pad_head_rows, resi_conv_specs, winattn_head_padded

RGB_MEAN = (0.4488, 0.4371, 0.4040)


def _lin_specs(prefix, cin, cout, bias=True):
    out = [(f'{prefix}.weight', (cout, cin), 'normal:0.06')]
    if bias:
        out.append((f'{prefix}.bias', (cout,), 'normal:0.02'))
    return out


def _ln_specs(prefix, c):
    return [(f'{prefix}.weight', (c,), 'affine_w'), (f'{prefix}.bias', (c,), 'normal:0.05')]


def _bn_specs(prefix, c):
    return [
        (f'{prefix}.weight', (c,), 'affine_w'),
        (f'{prefix}.bias', (c,), 'normal:0.05'),
        (f'{prefix}.running_mean', (c,), 'buffer_normal:0.1'),
        (f'{prefix}.running_var', (c,), 'buffer_var'),
        (f'{prefix}.num_batches_tracked', (), 'buffer_long'),
    ]


def _is_shifted(rg: int, b: int) -> bool:
    # arch.py:335 / :456
    return (rg % 2 == 0 and b > 0 and (b - 2) % 4 == 0) or (rg % 2 != 0 and b % 4 == 0)


def _rpe_buffers(hs: int, ws: int):
    """rpe_biases / relative_position_index exactly as Spatial_Attention.__init__ builds them (arch.py:193-213)."""
    bh, bw = torch.arange(1 - hs, hs), torch.arange(1 - ws, ws)
    biases = torch.stack(torch.meshgrid([bh, bw], indexing='ij')).flatten(1).transpose(0, 1).contiguous().float()
    coords = torch.stack(torch.meshgrid([torch.arange(hs), torch.arange(ws)], indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += hs - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return biases, rel.sum(-1)


def _shift_mask(hp: int, wp: int, hs: int, ws: int) -> torch.Tensor:
    """attn_mask buffer of a shifted block for an hp x wp input (arch.py:363-428); only stored, never read by the engine."""
    sh, sw = hs // 2, ws // 2
    img = torch.zeros(hp, wp)
    cnt = 0
    for a in (slice(0, -hs), slice(-hs, -sh), slice(-sh, None)):
        for b in (slice(0, -ws), slice(-ws, -sw), slice(-sw, None)):
            img[a, b] = cnt
            cnt += 1
    win = img.view(hp // hs, hs, wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, hs * ws)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class DAT(EngineModule):
    def __init__(
        self,
        img_size: int = 64,
        in_chans: int = 3,
        embed_dim: int = 180,
        split_size=(8, 32),
        depth=(6, 6, 6, 6, 6, 6),
        num_heads=(6, 6, 6, 6, 6, 6),
        expansion_factor: float = 2.0,
        qkv_bias: bool = True,
        upscale: int = 4,
        img_range: float = 1.0,
        resi_connection: str = '1conv',
        upsampler: str = 'pixelshuffle',
        seed: int = 0,
    ):
        if resi_connection not in ('1conv', '3conv') or upsampler not in ('pixelshuffle', 'pixelshuffledirect'):
            raise ValueError(f'unknown resi_connection {resi_connection!r} / upsampler {upsampler!r}')
        if upsampler == 'pixelshuffle' and upscale & (upscale - 1) and upscale != 3:
            raise ValueError(f'scale {upscale} is not supported. Supported scales: 2^n and 3.')  # Upsample, dat/arch.py:783-801
        dim, hidden = embed_dim, int(embed_dim * expansion_factor)
        split = [int(split_size[0]), int(split_size[1])]
        specs = conv_specs('conv_first', in_chans, dim, 3) + _ln_specs('before_RG.1', dim)
        for rg, (nblk, heads) in enumerate(zip(depth, num_heads)):
            if dim % heads or heads % 2:
                raise ValueError('embed_dim must be divisible by an even head count')
            for b in range(nblk):
                p = f'layers.{rg}.blocks.{b}'
                specs += _ln_specs(f'{p}.norm1', dim)
                if b % 2 == 1:
                    specs += [(f'{p}.attn.temperature', (heads, 1, 1), 'affine_w')]
                specs += _lin_specs(f'{p}.attn.qkv', dim, 3 * dim, qkv_bias) + _lin_specs(f'{p}.attn.proj', dim, dim)
                if b % 2 == 0:
                    pos_dim = (dim // 2 // 4) // 4
                    for br in (0, 1):
                        hs, ws = (split[0], split[1]) if br == 0 else (split[1], split[0])
                        biases, index = _rpe_buffers(hs, ws)
                        q = f'{p}.attn.attns.{br}'
                        specs += [(f'{q}.rpe_biases', biases, 'buffer_tensor'), (f'{q}.relative_position_index', index, 'buffer_tensor')]
                        specs += _lin_specs(f'{q}.pos.pos_proj', 2, pos_dim)
                        for name, out in (('pos1', pos_dim), ('pos2', pos_dim), ('pos3', heads // 2)):
                            specs += _ln_specs(f'{q}.pos.{name}.0', pos_dim) + _lin_specs(f'{q}.pos.{name}.2', pos_dim, out)
                    if _is_shifted(rg, b):
                        specs += [(f'{p}.attn.attn_mask_0', _shift_mask(img_size, img_size, split[0], split[1]), 'buffer_tensor'),
                                  (f'{p}.attn.attn_mask_1', _shift_mask(img_size, img_size, split[1], split[0]), 'buffer_tensor')]
                a = f'{p}.attn'
                specs += [(f'{a}.dwconv.0.weight', (dim, 1, 3, 3), 'conv_w'), (f'{a}.dwconv.0.bias', (dim,), 'bias:9')] + _bn_specs(f'{a}.dwconv.1', dim)
                specs += conv_specs(f'{a}.channel_interaction.1', dim, dim // 8, 1) + _bn_specs(f'{a}.channel_interaction.2', dim // 8)
                specs += conv_specs(f'{a}.channel_interaction.4', dim // 8, dim, 1)
                specs += conv_specs(f'{a}.spatial_interaction.0', dim, dim // 16, 1) + _bn_specs(f'{a}.spatial_interaction.1', dim // 16)
                specs += conv_specs(f'{a}.spatial_interaction.3', dim // 16, 1, 1)
                f = f'{p}.ffn'
                specs += _lin_specs(f'{f}.fc1', dim, hidden) + _ln_specs(f'{f}.sg.norm', hidden // 2)
                specs += [(f'{f}.sg.conv.weight', (hidden // 2, 1, 3, 3), 'conv_w'), (f'{f}.sg.conv.bias', (hidden // 2,), 'bias:9')]
                specs += _lin_specs(f'{f}.fc2', hidden // 2, dim) + _ln_specs(f'{p}.norm2', dim)
            specs += resi_conv_specs(f'layers.{rg}.conv', dim, resi_connection)
        specs += _ln_specs('norm', dim) + resi_conv_specs('conv_after_body', dim, resi_connection)
        if upsampler == 'pixelshuffle':
            specs += conv_specs('conv_before_upsample.0', dim, 64, 3)
            for i, r in enumerate([3] if upscale == 3 else [2] * int(math.log2(upscale))):
                specs += conv_specs(f'upsample.{2 * i}', 64, 64 * r * r, 3)
            specs += conv_specs('conv_last', 64, in_chans, 3)
        else:  # UpsampleOneStep (arch.py:804-825): one conv + PixelShuffle straight to the image
            specs += conv_specs('upsample.0', dim, upscale * upscale * in_chans, 3, gain=2.0)
        super().__init__(specs, in_chans, in_chans, upscale, seed=seed)
        self.resi_connection, self.upsampler_kind = resi_connection, upsampler
        self.dim, self.hidden, self.split, self.depth, self.heads = dim, hidden, split, list(depth), list(num_heads)
        self.img_range, self.img_size = float(img_range), img_size

    # ------------------------------------------------------------------ host-side weight algebra
    @staticmethod
    def _fold_bn(w, conv: str, bn: str):
        """eval-mode BatchNorm folded into the conv before it: y = (conv(x) - mean) / sqrt(var + eps) * g + b."""
        s = w[f'{bn}.weight'] / torch.sqrt(w[f'{bn}.running_var'] + 1e-5)
        weight = w[f'{conv}.weight'] * s.view(-1, *([1] * (w[f'{conv}.weight'].dim() - 1)))
        bias = (w[f'{conv}.bias'] - w[f'{bn}.running_mean']) * s + w[f'{bn}.bias']
        return weight, bias

    @staticmethod
    def _pos_table(w, q: str) -> torch.Tensor:
        """DynamicPosBias (residual=False, arch.py:135-143) over the rpe_biases mother set -> [offsets][heads/2]."""
        lin = lambda name, t: F.linear(t, w[f'{name}.weight'], w[f'{name}.bias'])
        t = lin(f'{q}.pos.pos_proj', w[f'{q}.rpe_biases'])
        for blk in ('pos1', 'pos2', 'pos3'):
            t = F.layer_norm(t, (t.shape[-1],), w[f'{q}.pos.{blk}.0.weight'], w[f'{q}.pos.{blk}.0.bias'], 1e-5)
            t = lin(f'{q}.pos.{blk}.2', F.relu(t))
        return t

    def build_plan(self, pb: PlanBuilder, w) -> None:
        dim, hidden = self.dim, self.hidden
        lin_w = lambda name: w[f'{name}.weight'].view(*w[f'{name}.weight'].shape, 1, 1)
        lin_b = lambda name: w.get(f'{name}.bias')
        feat, x, xn = pb.buffer(dim), pb.buffer(dim), pb.buffer(dim)
        # window blocks on a bf16 plan with head_dim < 32 and 8-aligned windows of 64 / 128 / 256 tokens: every head of q, k, v on a
        # 32-channel stride (zero rows of the qkv weights) -> the tcgen05 window-attention kernel (winattn_tc.cu).  Everything between
        # qkv and proj of such a block (attention output, depthwise-conv branch, AIM) then lives in that head-padded channel space:
        # the padding channels are zeros that meet zero weights, proj's padded input columns are zero
        padded = {h: winattn_head_padded(pb.compute_dtype, dim, h, self.split) for h in set(self.heads)}
        apad = max([dim] + [h * HEAD_PAD for h in padded if padded[h]])
        pad = max((dim + 15) // 16 * 16, apad)  # q | k | v (and the two FFN halves) start on 16-channel boundaries
        qkv = pb.buffer(3 * pad)
        att, convx, y = pb.buffer(apad), pb.buffer(apad), pb.buffer(apad)
        half = hidden // 2
        hpad = (half + 15) // 16 * 16
        hid, gate_n, gated = pb.buffer(2 * hpad), pb.buffer(half), pb.buffer(half)
        rg_res, img = pb.buffer(dim), pb.buffer(dim)
        stats = pb.buffer(8)  # LayerNorm statistics of norm1 / norm2, folded into the linears that consume them
        # hidden layer of AIM's spatial-interaction MLP (dim -> dim/16, GELU): a 1x1 conv on the tensor cores when the plan has them
        si_hid = pb.buffer(16) if pb.compute_dtype == torch.bfloat16 and dim // 16 <= 16 else None
        tmp_a = tmp_b = None
        if self.resi_connection == '3conv':
            tmp_a, tmp_b = pb.buffer(dim // 4), pb.buffer(dim // 4)
        mean = RGB_MEAN if self.in_channels == 3 else (0.0, 0.0, 0.0)
        pb.conv(INPUT, feat, w['conv_first.weight'], w['conv_first.bias'], in_mean=mean, in_scale=self.img_range)
        pb.layernorm(feat, x, w['before_RG.1.weight'], w['before_RG.1.bias'])
        for rg, (nblk, heads) in enumerate(zip(self.depth, self.heads)):
            # res = x (ResidualGroup keeps its input, arch.py:768): copy through an identity-free path: LN output already in x,
            # so stash it with a 1x1 identity conv only when the group has blocks that overwrite x
            pb.conv(x, rg_res, torch.eye(dim).view(dim, dim, 1, 1), None)
            # bf16 plan: `x += proj(..)` and `x += fc2(..)` also write the per-pixel {sum, sum of squares} of what they store
            # (rsb_conv_desc.ln_out): norm2 and the next block's norm1 need no statistics pass of their own
            fuse_stats = pb.ln_out_supported(dim)
            raw = False  # `stats` holds raw sums (written by a conv) instead of the statistics op's {rstd, -mean * rstd}
            for b in range(nblk):
                p = f'layers.{rg}.blocks.{b}'
                a = f'{p}.attn'
                # norm1 -> qkv and norm2 -> fc1: the normalised map is never written (one statistics pass, LayerNorm applied in the
                # linears' epilogues)
                if not raw:
                    pb.layernorm_stats(x, stats)
                ln1 = (stats, w[f'{p}.norm1.weight'], w[f'{p}.norm1.bias']) + ((1e-5,) if raw else ())
                wq, bq = lin_w(f'{a}.qkv'), lin_b(f'{a}.qkv')
                hp = b % 2 == 0 and padded[heads]
                width = heads * HEAD_PAD if hp else dim            # channels of this block's attention space
                rows_p = (lambda t: pad_head_rows(t, dim, heads)) if hp else (lambda t: t)   # output-channel axis -> padded layout
                cols_p = (lambda t: pad_head_cols(t, dim, heads)) if hp else (lambda t: t)   # input-channel axis
                for part in range(3):  # one conv per q / k / v (UMMA N <= 256)
                    rows = slice(part * dim, (part + 1) * dim)
                    pb.conv(x, qkv.slice(part * pad, width), rows_p(wq[rows]), None if bq is None else rows_p(bq[rows]), ln=ln1)
                att_b, convx_b, y_b = att.slice(0, width), convx.slice(0, width), y.slice(0, width)
                if b % 2 == 0:
                    t0, t1 = self._pos_table(w, f'{a}.attns.0'), self._pos_table(w, f'{a}.attns.1')
                    pb.op(N.OP_WINATTN, qkv, att_b, dim, ints=(heads, self.split[0], self.split[1], int(_is_shifted(rg, b)), pad, HEAD_PAD if hp else 0),
                          floats=((dim // heads) ** -0.5,), weights=(t0, t1))
                else:
                    pb.op(N.OP_CHANATTN, qkv, att_b, dim, ints=(heads, pad), weights=(w[f'{a}.temperature'],))
                dw_w, dw_b = self._fold_bn(w, f'{a}.dwconv.0', f'{a}.dwconv.1')
                pb.dwconv3(qkv.slice(2 * pad, width), convx_b, rows_p(dw_w), rows_p(dw_b), act=N.ACT_GELU)
                ci_w1, ci_b1 = self._fold_bn(w, f'{a}.channel_interaction.1', f'{a}.channel_interaction.2')
                si_w1, si_b1 = self._fold_bn(w, f'{a}.spatial_interaction.0', f'{a}.spatial_interaction.1')
                ci_w1, si_w1 = cols_p(ci_w1), cols_p(si_w1)
                hid_id = 0
                if si_hid is not None:  # spatial map source: the attention output in window blocks, the conv branch in channel blocks
                    pb.conv(att_b if b % 2 == 0 else convx_b, si_hid.slice(0, dim // 16), si_w1.reshape(dim // 16, width, 1, 1), si_b1, act=N.ACT_GELU)
                    hid_id = si_hid.buf + 1
                pb.op(N.OP_AIM, att_b, y_b, width, src2=convx_b, ints=(b % 2, dim // 8, dim // 16, hid_id),
                      weights=(ci_w1, ci_b1, rows_p(w[f'{a}.channel_interaction.4.weight']), rows_p(w[f'{a}.channel_interaction.4.bias']),
                               si_w1, si_b1, w[f'{a}.spatial_interaction.3.weight'], w[f'{a}.spatial_interaction.3.bias']))
                pb.conv(y_b, x, cols_p(lin_w(f'{a}.proj')), lin_b(f'{a}.proj'), combine=N.COMB_AXPY, res1=x,
                        ln_out=stats if fuse_stats else None)                                                          # x += proj(...)
                f = f'{p}.ffn'
                if not fuse_stats:
                    pb.layernorm_stats(x, stats)
                ln2 = (stats, w[f'{p}.norm2.weight'], w[f'{p}.norm2.bias']) + ((1e-5,) if fuse_stats else ())
                w1, b1 = lin_w(f'{f}.fc1'), lin_b(f'{f}.fc1')
                for part in range(2):  # x1 | x2 = chunk(2) of the hidden activations, each on its own plane range
                    rows = slice(part * half, (part + 1) * half)
                    pb.conv(x, hid.slice(part * hpad, half), w1[rows], b1[rows], act=N.ACT_GELU, ln=ln2)
                pb.layernorm(hid.slice(hpad, half), gate_n, w[f'{f}.sg.norm.weight'], w[f'{f}.sg.norm.bias'])
                pb.dwconv3(gate_n, gated, w[f'{f}.sg.conv.weight'], w[f'{f}.sg.conv.bias'], gate=hid.slice(0, half))
                raw = fuse_stats and b + 1 < nblk  # the next block's norm1 statistics
                pb.conv(gated, x, lin_w(f'{f}.fc2'), lin_b(f'{f}.fc2'), combine=N.COMB_AXPY, res1=x, ln_out=stats if raw else None)      # x += fc2(...)
            emit_resi_conv(pb, w, f'layers.{rg}.conv', self.resi_connection, x, img, rg_res, tmp_a, tmp_b)
            x, img = img, x
        pb.layernorm(x, xn, w['norm.weight'], w['norm.bias'])
        y = y.slice(0, dim)
        emit_resi_conv(pb, w, 'conv_after_body', self.resi_connection, xn, y, feat, tmp_a, tmp_b)
        omean = mean if self.in_channels == 3 else (0.0, 0.0, 0.0)
        if self.upsampler_kind == 'pixelshuffledirect':
            pb.conv(y, OUTPUT, w['upsample.0.weight'], w['upsample.0.bias'], ps=self.upscale, out_scale=1.0 / self.img_range, out_mean=omean)
            return
        cur = pb.buffer(64)
        pb.conv(y, cur, w['conv_before_upsample.0.weight'], w['conv_before_upsample.0.bias'], act=N.ACT_LRELU, act_param=0.01)
        grid = 1
        for i, r in enumerate([3] if self.upscale == 3 else [2] * int(math.log2(self.upscale))):
            nxt = pb.buffer(64, scale=grid * r)
            # conv 64 -> 64 r^2 + PixelShuffle(r) (Upsample, arch.py:783-801: 2^n as n x2 steps, 3 as one x3 step):
            # channel c * r^2 + phase -> one conv per sub-pixel phase, stored pixel-interleaved
            wk, bk = w[f'upsample.{2 * i}.weight'], w[f'upsample.{2 * i}.bias']
            perm = torch.arange(64 * r * r).view(64, r * r).t().reshape(-1)
            for phase in range(r * r):
                sel = perm[phase * 64:(phase + 1) * 64]
                pb.conv(cur, nxt, wk[sel], bk[sel], dst_ps=r, dst_phase=phase)
            cur, grid = nxt, grid * r
        pb.conv(cur, OUTPUT, w['conv_last.weight'], w['conv_last.bias'], ps=1, out_scale=1.0 / self.img_range, out_mean=omean)


class DatArch(Architecture[DAT]):
    def __init__(self):
        super().__init__(
            uid='dat',
            detect=KeyCondition.has_all(
                'conv_first.weight', 'before_RG.1.weight', 'before_RG.1.bias',
                'layers.0.blocks.0.norm1.weight', 'layers.0.blocks.0.norm2.weight',
                'layers.0.blocks.0.ffn.fc1.weight', 'layers.0.blocks.0.ffn.sg.norm.weight',
                'layers.0.blocks.0.ffn.sg.conv.weight', 'layers.0.blocks.0.ffn.fc2.weight',
                'layers.0.blocks.0.attn.qkv.weight', 'layers.0.blocks.0.attn.proj.weight',
                'layers.0.blocks.0.attn.dwconv.0.weight', 'layers.0.blocks.0.attn.dwconv.1.running_mean',
                'layers.0.blocks.0.attn.channel_interaction.1.weight', 'layers.0.blocks.0.attn.channel_interaction.2.running_mean',
                'layers.0.blocks.0.attn.channel_interaction.4.weight', 'layers.0.blocks.0.attn.spatial_interaction.0.weight',
                'layers.0.blocks.0.attn.spatial_interaction.1.running_mean', 'layers.0.blocks.0.attn.spatial_interaction.3.weight',
                'layers.0.blocks.0.attn.attns.0.rpe_biases', 'layers.0.blocks.0.attn.attns.0.relative_position_index',
                'layers.0.blocks.0.attn.attns.0.pos.pos_proj.weight', 'layers.0.blocks.0.attn.attns.0.pos.pos1.0.weight',
                'layers.0.blocks.0.attn.attns.0.pos.pos3.0.weight', 'norm.weight',
            ),
        )

    def load(self, state_dict: Mapping[str, object]):
        in_chans, embed_dim = state_dict['conv_first.weight'].shape[1], state_dict['conv_first.weight'].shape[0]
        num_layers = get_seq_len(state_dict, 'layers')
        depth = [get_seq_len(state_dict, f'layers.{i}.blocks') for i in range(num_layers)]
        num_heads: List[int] = []
        for i in range(num_layers):
            if depth[i] >= 2:
                num_heads.append(state_dict[f'layers.{i}.blocks.1.attn.temperature'].shape[0])
            else:  # only even head counts are recoverable (dat/__init__.py:60-62)
                num_heads.append(state_dict[f'layers.{i}.blocks.0.attn.attns.0.pos.pos3.2.weight'].shape[0] * 2)
        upsampler = 'pixelshuffle' if 'conv_last.weight' in state_dict else 'pixelshuffledirect'
        resi_connection = '1conv' if 'conv_after_body.weight' in state_dict else '3conv'
        if upsampler == 'pixelshuffle':
            upscale = 1
            for i in range(0, get_seq_len(state_dict, 'upsample'), 2):
                wt = state_dict[f'upsample.{i}.weight']
                upscale *= int(math.sqrt(wt.shape[0] // wt.shape[1]))
        else:
            upscale = pixelshuffle_scale(state_dict['upsample.0.weight'].shape[0], in_chans)
        img_size = 64
        if 'layers.0.blocks.2.attn.attn_mask_0' in state_dict:
            mx, my, _ = state_dict['layers.0.blocks.2.attn.attn_mask_0'].shape
            img_size = int(math.sqrt(mx * my))
        split_size = [int(v) + 1 for v in state_dict['layers.0.blocks.0.attn.attns.0.rpe_biases'][-1]]
        model = DAT(
            img_size=img_size, in_chans=in_chans, embed_dim=embed_dim, split_size=split_size, depth=depth, num_heads=num_heads,
            expansion_factor=float(state_dict['layers.0.blocks.0.ffn.fc1.weight'].shape[0] / embed_dim),
            qkv_bias='layers.0.blocks.0.attn.qkv.bias' in state_dict, upscale=upscale, resi_connection=resi_connection, upsampler=upsampler,
        )
        return self._enhance_model(model=model, in_channels=in_chans, out_channels=in_chans, upscale=upscale, name='DAT')
