"""Op-level parity of the token / attention ops (rsb_plan_add_op) against fp64 restatements of the reference math, through the C ABI.

VERDICT r1 ("What's weak" 2): whole-model PSNR is a loose bar for one op buried in 36 blocks; here every op is alone between two
identity convs.  References follow /root/reference/resselt/archs/dat/arch.py:224-267 (window attention incl. dynamic position
bias lookup, shift + mask :363-428,456-482, zero padding :443-449), :565-589 (channel attention), :40-101 / :345-352 (depthwise
convs), :492-508 / :594-607 (AIM), swinir/arch.py:133-170,268-293 (W-MSA / SW-MSA), plksr/rplksr.py:83,91-93 (GroupNorm + skip),
utilities/dysample.py:46-83.
bf16 plan: inputs are rounded to bf16 first, the fp64 reference runs on the rounded values, the result may differ by the op's own
bf16 roundings; fp32 plan (CUDA-core kernels): <= 1e-4 range-normalised."""
import math

import pytest
import torch
import torch.nn.functional as F

from resselt_b200.archs._common import dysample_init_pos, emit_dysample
from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
DTYPES = [torch.bfloat16, torch.float32]


def _q(t, dtype):
    """What the engine sees of a host tensor: bf16-rounded on the bf16 plan."""
    return t.to(dtype).double()


def _check(got, ref, dtype, bf16_tol=1.2e-2, what=''):
    span = max(1.0, float(ref.max() - ref.min()))
    err = float((got.double() - ref).abs().max()) / span
    tol = bf16_tol if dtype == torch.bfloat16 else 1e-4
    assert math.isfinite(err) and err <= tol, f'{what}: range-normalised max-abs error {err:.3e} > {tol:.1e}'
    assert float(ref.abs().max()) > 1e-3, 'vacuous reference'


def _select(rows, cols, start):
    """[rows x cols] 0/1 matrix picking input channels start .. start+rows-1 (a 1x1 conv that moves a channel range)."""
    w = torch.zeros(rows, cols, 1, 1)
    w[torch.arange(rows), start + torch.arange(rows)] = 1.0
    return w


def _run(pb, x, dtype):
    plan = pb.finalize(torch.device(DEV))
    y = plan.forward(x.to(DEV, dtype)).double().cpu()
    torch.cuda.synchronize()
    return y, plan


# ------------------------------------------------------------------------------------------------ window attention
def _shift_mask(Hp, Wp, Hs, Ws, sh, sw):
    img = torch.zeros(Hp, Wp, dtype=torch.float64)
    cnt = 0
    for hs in (slice(0, -Hs), slice(-Hs, -sh), slice(-sh, None)):
        for ws in (slice(0, -Ws), slice(-Ws, -sw), slice(-sw, None)):
            img[hs, ws] = cnt
            cnt += 1
    win = img.view(Hp // Hs, Hs, Wp // Ws, Ws).permute(0, 2, 1, 3).reshape(-1, Hs * Ws)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _rel_index(Hs, Ws):
    """relative_position_index of the reference: (y_query - y_key + Hs - 1) * (2 Ws - 1) + (x_query - x_key + Ws - 1)."""
    ys, xs = torch.meshgrid(torch.arange(Hs), torch.arange(Ws), indexing='ij')
    c = torch.stack([ys.flatten(), xs.flatten()])
    rel = c[:, :, None] - c[:, None, :]
    return (rel[0] + Hs - 1) * (2 * Ws - 1) + rel[1] + Ws - 1


def ref_window_attention(q, k, v, heads, split, shifted, scale, tables):
    """q, k, v: [B, C, H, W] fp64.  Two branches on the two channel halves: branch 0 windows split[0] x split[1], branch 1 transposed."""
    B, C, H, W = q.shape
    m = max(split)
    Hp, Wp = (H + m - 1) // m * m, (W + m - 1) // m * m
    pad = lambda t: F.pad(t, (0, Wp - W, 0, Hp - H)).permute(0, 2, 3, 1)  # zero padding of q, k, v themselves (arch.py:443-449)
    q, k, v = pad(q), pad(k), pad(v)
    half, hb = C // 2, heads // 2
    d = half // hb
    outs = []
    for br in (0, 1):
        Hs, Ws = (split[0], split[1]) if br == 0 else (split[1], split[0])
        sh, sw = Hs // 2, Ws // 2
        sl = slice(br * half, (br + 1) * half)
        tq, tk, tv = q[..., sl], k[..., sl], v[..., sl]
        mask = None
        if shifted:
            tq, tk, tv = (torch.roll(t, shifts=(-sh, -sw), dims=(1, 2)) for t in (tq, tk, tv))
            mask = _shift_mask(Hp, Wp, Hs, Ws, sh, sw)

        def win(t):
            t = t.reshape(B, Hp // Hs, Hs, Wp // Ws, Ws, half).permute(0, 1, 3, 2, 4, 5).reshape(-1, Hs * Ws, hb, d)
            return t.permute(0, 2, 1, 3)

        attn = (win(tq) * scale) @ win(tk).transpose(-2, -1)
        bias = tables[br].double()[_rel_index(Hs, Ws).view(-1)].view(Hs * Ws, Hs * Ws, hb).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if mask is not None:
            nW = mask.shape[0]
            attn = (attn.view(B, nW, hb, Hs * Ws, Hs * Ws) + mask.view(1, nW, 1, Hs * Ws, Hs * Ws)).view(-1, hb, Hs * Ws, Hs * Ws)
        o = (attn.softmax(-1) @ win(tv)).transpose(1, 2).reshape(-1, Hs * Ws, half)
        o = o.view(B, Hp // Hs, Wp // Ws, Hs, Ws, half).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, half)
        if shifted:
            o = torch.roll(o, shifts=(sh, sw), dims=(1, 2))
        outs.append(o[:, :H, :W])
    return torch.cat(outs, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('dim,heads,split,shifted,B,H,W', [
    (60, 2, (8, 32), 0, 1, 64, 64),      # DAT branch shapes: 8x32 / 32x8 windows, head_dim 30
    (60, 2, (8, 32), 1, 1, 64, 96),      # shifted: roll + region mask
    (60, 2, (8, 32), 1, 2, 40, 72),      # H and W padded up to multiples of 32 (zero q/k/v tokens take part in the softmax)
    (120, 4, (32, 8), 1, 1, 50, 33),     # transposed split first, two heads per branch, both dims padded
    (64, 2, (8, 8), 0, 1, 48, 64),       # Swin: square windows, head_dim 32
    (60, 2, (8, 8), 1, 2, 32, 40),       # Swin shifted
    (56, 4, (7, 7), 1, 1, 49, 35),       # window 7 (classical SwinIR), head_dim 14
    (48, 6, (4, 16), 0, 1, 32, 32),      # head_dim 8, three heads per branch
])
def test_window_attention_op(dim, heads, split, shifted, B, H, W, dtype):
    g = torch.Generator().manual_seed(dim * 31 + H * 7 + W + shifted)
    pad = (dim + 15) // 16 * 16
    x = torch.randn(B, 3 * dim, H, W, generator=g)
    hb = heads // 2
    tabs = [torch.randn((2 * split[br] - 1) * (2 * split[1 - br] - 1), hb, generator=g) * 0.5 for br in (0, 1)]
    scale = (dim // heads) ** -0.5
    pb = PlanBuilder(dtype, 3 * dim, dim, 1)
    raw, qkv, att = pb.buffer(3 * dim), pb.buffer(3 * pad), pb.buffer(dim)
    pb.conv(INPUT, raw, torch.eye(3 * dim).view(3 * dim, 3 * dim, 1, 1))
    for part in range(3):
        pb.conv(raw, qkv.slice(part * pad, dim), _select(dim, 3 * dim, part * dim))
    pb.op(N.OP_WINATTN, qkv, att, dim, ints=(heads, split[0], split[1], shifted, pad), floats=(scale,), weights=(tabs[0], tabs[1]))
    pb.conv(att, OUTPUT, torch.eye(dim).view(dim, dim, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    ref = ref_window_attention(xq[:, :dim], xq[:, dim:2 * dim], xq[:, 2 * dim:], heads, split, shifted, scale, tabs)
    # bf16 kernel: P is rounded to bf16 before PV and the output once more
    _check(got, ref, dtype, bf16_tol=1.5e-2, what=f'window attention {split} shifted={shifted}')


@pytest.mark.parametrize('dim,heads,split,shifted,B,H,W', [
    (60, 2, (8, 32), 0, 1, 64, 64),      # DAT: 256-token windows = two 128-query tiles per window
    (60, 2, (8, 32), 1, 1, 64, 96),      # shifted: roll + region mask in the seam windows
    (180, 6, (8, 32), 1, 2, 40, 72),     # DAT's real head layout (3 heads per branch), H and W padded (zero tokens take part)
    (120, 4, (32, 8), 1, 1, 50, 33),     # transposed split first
    (60, 2, (8, 8), 0, 1, 48, 64),       # Swin: two 64-token windows per tile (block-diagonal P)
    (180, 6, (8, 8), 1, 2, 40, 56),      # Swin shifted, odd number of windows (ragged last tile)
    (48, 2, (8, 16), 1, 1, 32, 48),      # 128-token windows, head_dim 24
    (32, 4, (16, 16), 0, 1, 32, 64),     # 16x16 windows, head_dim 8
])
def test_window_attention_tc_op(dim, heads, split, shifted, B, H, W):
    """Head-padded layout -> tcgen05 kernel (csrc/winattn_tc.cu): heads on 32-channel strides in q / k / v and in the output."""
    from resselt_b200.archs._common import HEAD_PAD, head_pad_index, winattn_head_padded
    dtype = torch.bfloat16
    assert winattn_head_padded(dtype, dim, heads, split)
    g = torch.Generator().manual_seed(dim * 31 + H * 7 + W + shifted)
    pad = heads * HEAD_PAD
    x = torch.randn(B, 3 * dim, H, W, generator=g)
    hb = heads // 2
    tabs = [torch.randn((2 * split[br] - 1) * (2 * split[1 - br] - 1), hb, generator=g) * 0.5 for br in (0, 1)]
    scale = (dim // heads) ** -0.5
    pb = PlanBuilder(dtype, 3 * dim, dim, 1)
    raw, qkv, att = pb.buffer(3 * dim), pb.buffer(3 * pad), pb.buffer(pad)
    pb.conv(INPUT, raw, torch.eye(3 * dim).view(3 * dim, 3 * dim, 1, 1))
    idx = head_pad_index(dim, heads)
    for part in range(3):
        sel = torch.zeros(pad, 3 * dim, 1, 1)
        sel[idx, part * dim + torch.arange(dim)] = 1.0
        pb.conv(raw, qkv.slice(part * pad, pad), sel)
    pb.op(N.OP_WINATTN, qkv, att, dim, ints=(heads, split[0], split[1], shifted, pad, HEAD_PAD), floats=(scale,), weights=(tabs[0], tabs[1]))
    back = torch.zeros(dim, pad, 1, 1)
    back[torch.arange(dim), idx] = 1.0
    pb.conv(att, OUTPUT, back)
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    ref = ref_window_attention(xq[:, :dim], xq[:, dim:2 * dim], xq[:, 2 * dim:], heads, split, shifted, scale, tabs)
    _check(got, ref, dtype, bf16_tol=1.5e-2, what=f'tcgen05 window attention {split} shifted={shifted}')
    # the padded channels of the output are exact zeros (they feed zero weight columns, but must not be NaN / Inf)
    pb2 = PlanBuilder(dtype, 3 * dim, pad, 1)
    raw, qkv, att = pb2.buffer(3 * dim), pb2.buffer(3 * pad), pb2.buffer(pad)
    pb2.conv(INPUT, raw, torch.eye(3 * dim).view(3 * dim, 3 * dim, 1, 1))
    for part in range(3):
        sel = torch.zeros(pad, 3 * dim, 1, 1)
        sel[idx, part * dim + torch.arange(dim)] = 1.0
        pb2.conv(raw, qkv.slice(part * pad, pad), sel)
    pb2.op(N.OP_WINATTN, qkv, att, dim, ints=(heads, split[0], split[1], shifted, pad, HEAD_PAD), floats=(scale,), weights=(tabs[0], tabs[1]))
    pb2.conv(att, OUTPUT, torch.eye(pad).view(pad, pad, 1, 1))
    full, _ = _run(pb2, x, dtype)
    keep = torch.zeros(pad, dtype=torch.bool)
    keep[idx] = True
    assert float(full[:, ~keep].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ channel attention
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('dim,heads,B,H,W', [(60, 2, 1, 48, 40), (180, 6, 1, 64, 72), (64, 2, 2, 33, 17), (96, 4, 1, 128, 128)])
def test_channel_attention_op(dim, heads, B, H, W, dtype):
    g = torch.Generator().manual_seed(dim + H)
    pad = (dim + 15) // 16 * 16
    x = torch.randn(B, 3 * dim, H, W, generator=g)
    temp = 1.0 + 0.3 * torch.rand(heads, 1, 1, generator=g)
    pb = PlanBuilder(dtype, 3 * dim, dim, 1)
    raw, qkv, att = pb.buffer(3 * dim), pb.buffer(3 * pad), pb.buffer(dim)
    pb.conv(INPUT, raw, torch.eye(3 * dim).view(3 * dim, 3 * dim, 1, 1))
    for part in range(3):
        pb.conv(raw, qkv.slice(part * pad, dim), _select(dim, 3 * dim, part * dim))
    pb.op(N.OP_CHANATTN, qkv, att, dim, ints=(heads, pad), weights=(temp,))
    pb.conv(att, OUTPUT, torch.eye(dim).view(dim, dim, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    d = dim // heads
    q, k, v = (xq[:, i * dim:(i + 1) * dim].reshape(B, heads, d, H * W) for i in range(3))
    attn = (F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)) * temp.double().view(1, heads, 1, 1)
    ref = (attn.softmax(-1) @ v).reshape(B, dim, H, W)
    _check(got, ref, dtype, what='channel attention')


# ------------------------------------------------------------------------------------------------ depthwise 3x3
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,act,gated,B,H,W', [(180, N.ACT_GELU, False, 1, 37, 50), (184, N.ACT_NONE, True, 1, 40, 33), (64, N.ACT_GELU, False, 2, 9, 130),
                                              (96, N.ACT_NONE, True, 1, 128, 128)])
def test_depthwise3x3_op(C, act, gated, B, H, W, dtype):
    g = torch.Generator().manual_seed(C + W)
    cin = 2 * C if gated else C
    x = torch.randn(B, cin, H, W, generator=g)
    wt, bias = torch.randn(C, 1, 3, 3, generator=g) / 3.0, torch.randn(C, generator=g) * 0.2
    pb = PlanBuilder(dtype, cin, C, 1)
    raw, dst = pb.buffer(cin), pb.buffer(C)
    pb.conv(INPUT, raw, torch.eye(cin).view(cin, cin, 1, 1))
    pb.dwconv3(raw.slice(0, C), dst, wt, bias, act=act, gate=raw.slice(C, C) if gated else None)
    pb.conv(dst, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    ref = F.conv2d(xq[:, :C], wt.double(), bias.double(), padding=1, groups=C)
    if act == N.ACT_GELU:
        ref = F.gelu(ref)
    if gated:
        ref = ref * xq[:, C:]
    _check(got, ref, dtype, what='depthwise 3x3')


# ------------------------------------------------------------------------------------------------ AIM (adaptive interaction module)
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('hid_conv', [False, True])
@pytest.mark.parametrize('mode,C,B,H,W', [(0, 180, 1, 48, 56), (1, 180, 1, 48, 56), (0, 64, 2, 31, 20), (1, 96, 1, 128, 128)])
def test_aim_op(mode, C, B, H, W, hid_conv, dtype):
    """hid_conv: the spatial MLP's hidden layer comes from a 1x1 conv op (tensor cores) instead of the AIM kernel itself (i[3], bf16 plans)."""
    if hid_conv and dtype != torch.bfloat16:
        pytest.skip('bf16-plan feature')
    g = torch.Generator().manual_seed(C + mode)
    h1, h2 = C // 8, C // 16
    x = torch.randn(B, 2 * C, H, W, generator=g)  # [attention output | conv branch]
    r = lambda *s: torch.randn(*s, generator=g)
    ci_w1, ci_b1, ci_w2, ci_b2 = r(h1, C) / C ** 0.5, r(h1) * 0.1, r(C, h1) / h1 ** 0.5, r(C) * 0.1
    si_w1, si_b1, si_w2, si_b2 = r(h2, C) / C ** 0.5, r(h2) * 0.1, r(1, h2) / h2 ** 0.5, r(1) * 0.1
    pb = PlanBuilder(dtype, 2 * C, C, 1)
    raw, att, convx, y = pb.buffer(2 * C), pb.buffer(C), pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, raw, torch.eye(2 * C).view(2 * C, 2 * C, 1, 1))
    pb.conv(raw, att, _select(C, 2 * C, 0))
    pb.conv(raw, convx, _select(C, 2 * C, C))
    hid_id = 0
    if hid_conv:
        hid = pb.buffer(16)
        pb.conv(att if mode == 0 else convx, hid.slice(0, h2), si_w1.view(h2, C, 1, 1), si_b1, act=N.ACT_GELU)
        hid_id = hid.buf + 1
    pb.op(N.OP_AIM, att, y, C, src2=convx, ints=(mode, h1, h2, hid_id), weights=(ci_w1, ci_b1, ci_w2, ci_b2, si_w1, si_b1, si_w2, si_b2))
    pb.conv(y, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    a, c = xq[:, :C], xq[:, C:]
    D = lambda t: t.double()

    def channel_map(t):   # [B, C, 1, 1]: 1x1 conv -> (folded BN) -> GELU -> 1x1 conv on the global average (arch.py:348-353)
        p = t.mean(dim=(2, 3))
        return (F.gelu(p @ D(ci_w1).t() + D(ci_b1)) @ D(ci_w2).t() + D(ci_b2)).view(B, C, 1, 1)

    def spatial_map(t):   # [B, 1, H, W] (arch.py:355-360)
        hid = F.gelu(torch.einsum('bchw,kc->bkhw', t, D(si_w1)) + D(si_b1).view(1, -1, 1, 1))
        return torch.einsum('bkhw,ok->bohw', hid, D(si_w2)) + D(si_b2).view(1, 1, 1, 1)

    if mode == 0:   # window-attention block (arch.py:492-508): channel map from the conv branch gates the attention, spatial map from the attention gates the conv branch
        ref = a * torch.sigmoid(channel_map(c)) + torch.sigmoid(spatial_map(a)) * c
    else:           # channel-attention block (arch.py:594-607)
        ref = a * torch.sigmoid(spatial_map(c)) + c * torch.sigmoid(channel_map(a))
    _check(got, ref, dtype, what=f'AIM mode {mode}')


# ------------------------------------------------------------------------------------------------ GroupNorm + skip
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,groups,B,H,W,skip', [(64, 4, 1, 256, 256, True), (64, 4, 2, 300, 177, True), (32, 2, 1, 40, 24, False), (64, 4, 1, 512, 512, True)])
def test_groupnorm_op(C, groups, B, H, W, skip, dtype):
    """RealPLKSR's GroupNorm(4, 64) + block skip at real map sizes: 16 channels x 262 144 pixels per group at 512^2 —
    the global reduction whose accuracy the small whole-model fixtures cannot show."""
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(B, 2 * C, H, W, generator=g) * 1.5 + 0.7   # non-zero mean: a one-pass variance would lose digits here
    gamma, beta = 1.0 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    pb = PlanBuilder(dtype, 2 * C, C, 1)
    raw, a, s, y = pb.buffer(2 * C), pb.buffer(C), pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, raw, torch.eye(2 * C).view(2 * C, 2 * C, 1, 1))
    pb.conv(raw, a, _select(C, 2 * C, 0))
    pb.conv(raw, s, _select(C, 2 * C, C))
    pb.groupnorm(a, y, groups, gamma, beta, eps=1e-5, skip=s if skip else None)
    pb.conv(y, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    ref = F.group_norm(xq[:, :C], groups, gamma.double(), beta.double(), 1e-5)
    if skip:
        ref = ref + xq[:, C:]
    _check(got, ref, dtype, bf16_tol=6e-3, what='GroupNorm')


# ------------------------------------------------------------------------------------------------ DySample head
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,out_ch,scale,groups,B,H,W', [(48, 3, 2, 4, 1, 40, 56), (64, 3, 4, 4, 1, 24, 33), (48, 3, 3, 3, 2, 17, 20)])
def test_dysample_op(C, out_ch, scale, groups, B, H, W, dtype):
    """DySample.forward (utilities/dysample.py:46-83) restated with grid_sample in fp64 on the same (bf16-rounded) features."""
    g = torch.Generator().manual_seed(C + scale)
    k = 2 * groups * scale * scale
    x = torch.randn(B, C, H, W, generator=g)
    w = {
        'up.end_conv.weight': torch.randn(out_ch, C, 1, 1, generator=g) / C ** 0.5, 'up.end_conv.bias': torch.randn(out_ch, generator=g) * 0.1,
        'up.offset.weight': torch.randn(k, C, 1, 1, generator=g) * 0.05, 'up.offset.bias': torch.randn(k, generator=g) * 0.05,
        'up.scope.weight': torch.randn(k, C, 1, 1, generator=g) * 0.05, 'up.init_pos': dysample_init_pos(scale, groups),
    }
    pb = PlanBuilder(dtype, C, out_ch, scale)
    feat = pb.buffer(C)
    pb.conv(INPUT, feat, torch.eye(C).view(C, C, 1, 1))
    emit_dysample(pb, {k_: v.double() for k_, v in w.items()}, 'up', feat, out_ch, scale, groups)
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    D = lambda name: w[name].double()
    offset = F.conv2d(xq, D('up.offset.weight'), D('up.offset.bias')) * torch.sigmoid(F.conv2d(xq, D('up.scope.weight'))) * 0.5 + D('up.init_pos')
    offset = offset.view(B, 2, -1, H, W)
    coords_h, coords_w = torch.arange(H, dtype=torch.float64) + 0.5, torch.arange(W, dtype=torch.float64) + 0.5
    coords = torch.stack(torch.meshgrid([coords_w, coords_h], indexing='ij')).transpose(1, 2).unsqueeze(1).unsqueeze(0)
    normalizer = torch.tensor([W, H], dtype=torch.float64).view(1, 2, 1, 1, 1)
    coords = 2 * (coords + offset) / normalizer - 1
    coords = F.pixel_shuffle(coords.reshape(B, -1, H, W), scale).view(B, 2, -1, scale * H, scale * W).permute(0, 2, 3, 4, 1).contiguous().flatten(0, 1)
    samp = F.grid_sample(xq.reshape(B * groups, -1, H, W), coords, mode='bilinear', align_corners=False, padding_mode='border')
    ref = F.conv2d(samp.view(B, -1, scale * H, scale * W), D('up.end_conv.weight'), D('up.end_conv.bias'))
    # bf16 plan: the offsets themselves are stored in bf16 before sampling (8 mantissa bits of a sub-pixel position)
    _check(got, ref, dtype, bf16_tol=3e-2, what='DySample')


# ------------------------------------------------------------------------------------------------ RTMoSR ops (rsb_op_kind 7-9, depthwise 5x5)
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,B,H,W', [(32, 1, 40, 56), (48, 2, 17, 30), (64, 1, 128, 96)])
def test_rmsnorm_op(C, B, H, W, dtype):
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    scale, offset = 1.0 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    pb = PlanBuilder(dtype, C, C, 1)
    a, b = pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.rmsnorm(a, b, scale, offset, eps=1e-6)
    pb.conv(b, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    rms = xq.norm(2, dim=1, keepdim=True) * C ** -0.5
    ref = scale.double().view(1, -1, 1, 1) * (xq / (rms + 1e-6)) + offset.double().view(1, -1, 1, 1)
    _check(got, ref, dtype, what='RMSNorm')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,cout,B,H,W', [(32, 64, 1, 40, 56), (256, 768, 1, 24, 40), (48, 72, 2, 17, 30)])
def test_rmsnorm_folded_into_linear(C, cout, B, H, W, dtype):
    """conv(x, ln=(rms stats, scale, offset)) == Conv1x1(RMSNorm(x)): statistics op in RMS mode (rsb_op_desc.i[0] = 2) + ln_fold."""
    g = torch.Generator().manual_seed(C + cout)
    x = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    scale, offset = 1.0 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    wl, bl = torch.randn(cout, C, 1, 1, generator=g) / C ** 0.5, 0.1 * torch.randn(cout, generator=g)
    pb = PlanBuilder(dtype, C, cout, 1)
    a, stats, y = pb.buffer(C), pb.buffer(8), pb.buffer(cout)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.rmsnorm_stats(a, stats, eps=1e-6)
    pb.conv(a, y, wl, bl, ln=(stats, scale, offset), act=N.ACT_MISH)
    pb.conv(y, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    rms = xq.norm(2, dim=1, keepdim=True) * C ** -0.5
    t = scale.double().view(1, -1, 1, 1) * (xq / (rms + 1e-6)) + offset.double().view(1, -1, 1, 1)
    ref = F.mish(F.conv2d(t, wl.double(), bl.double()))
    _check(got, ref, dtype, what='linear(RMSNorm(.))')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,se,B,H,W', [(32, True, 1, 40, 56), (48, False, 2, 18, 30), (32, True, 1, 128, 96)])
def test_unshuffle_pool_dw5_se_shuffle_chain(C, se, B, H, W, dtype):
    """GatedCNNBlock.conv of RTMoSR in isolation (rtmosr/arch.py:314-319) on a base-divisor-2 plan:
    PixelUnshuffle(2) + RepConv(MaxPool2d(2)) -> depthwise 5x5 -> [CSELayer] -> PixelShuffle(2)."""
    g = torch.Generator().manual_seed(C + W)
    x = torch.randn(B, C, H, W, generator=g)
    wp, bp = torch.randn(4 * C, C, 3, 3, generator=g) / (C * 9) ** 0.5, torch.randn(4 * C, generator=g) * 0.1
    w5, b5 = torch.randn(4 * C, 1, 5, 5, generator=g) / 5.0, torch.randn(4 * C, generator=g) * 0.1
    s1w, s1b = torch.randn(2 * C, 4 * C, 1, 1, generator=g) / (4 * C) ** 0.5 * 3, torch.randn(2 * C, generator=g) * 0.3
    s2w, s2b = torch.randn(4 * C, 2 * C, 1, 1, generator=g) / (2 * C) ** 0.5 * 3, torch.randn(4 * C, generator=g) * 0.3
    pb = PlanBuilder(dtype, C, C, 1, base_divisor=2)
    a, out = pb.buffer(C), pb.buffer(C)
    u5, v, o = pb.buffer(5 * C, scale=1), pb.buffer(4 * C, scale=1), pb.buffer(4 * C, scale=1)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.unshuffle_pool(a, u5)
    pb.conv(u5.slice(4 * C, C), v, wp, bp, combine=N.COMB_AXPY, res1=u5.slice(0, 4 * C))
    pb.dwconv(v, o, w5, b5)
    pb.se_shuffle(o, out, se=(s1w, s1b, s2w, s2b) if se else None)
    pb.conv(out, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, plan = _run(pb, x, dtype)
    xq = _q(x, dtype)
    D = lambda t: t.double()
    q = lambda t: t.to(dtype).double()  # a buffer between two ops holds bf16 on the bf16 plan
    pu, pool = F.pixel_unshuffle(xq, 2), F.max_pool2d(xq, 2, 2)
    assert torch.equal(plan.read_buffer(u5.slice(0, 4 * C)).double().cpu(), pu), 'PixelUnshuffle(2) is pure data movement'
    assert torch.equal(plan.read_buffer(u5.slice(4 * C, C)).double().cpu(), pool), 'MaxPool2d(2) is exact'
    t = q(pu + F.conv2d(pool, q(wp) if dtype == torch.bfloat16 else D(wp), D(bp), padding=1))
    t = q(F.conv2d(t, D(w5), D(b5), padding=2, groups=4 * C))
    if se:
        sq = t.mean(dim=(2, 3), keepdim=True)
        t = t * F.hardsigmoid(F.conv2d(F.relu(F.conv2d(sq, D(s1w), D(s1b))), D(s2w), D(s2b)))
    ref = F.pixel_shuffle(t, 2)
    _check(got, ref, dtype, bf16_tol=2e-2, what='unshuffle/pool -> dw5x5 -> SE -> shuffle')


@pytest.mark.parametrize('C,cin,cout,mean,B,H,W', [
    (180, 192, 180, 0.3, 1, 40, 56),     # SwinIR / DAT: x += proj(att) writes the sums norm2 -> fc1 consumes (two epilogue warpgroups per pixel)
    (180, 384, 360, 3.0, 2, 33, 47),     # x += fc2(hidden): K-chunked producer; token mean ten times the spread; image size not a multiple of 8
    (60, 64, 120, -1.0, 1, 64, 64),      # light models: one epilogue warpgroup per pixel (second pair of the chunk written as zeros)
    (180, 192, 180, 0.2, 1, 256, 256),   # many tiles
])
def test_layernorm_sums_from_the_producing_linear(C, cin, cout, mean, B, H, W):
    """rsb_conv_desc.ln_out / ln_fold = 2: the residual linear writes {sum, sum of squares} of what it stores, the next linear derives
    mean / rstd from them in its epilogue — Linear(LayerNorm(res + Linear(x))) against fp64, and against the statistics-op form."""
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(C + cin + H)
    x = torch.randn(B, cin + C, H, W, generator=g) * 0.5
    x[:, cin:] = x[:, cin:] * 0.8 + mean
    w1, b1 = torch.randn(C, cin, 1, 1, generator=g) / cin ** 0.5, 0.1 * torch.randn(C, generator=g)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    w2, b2 = torch.randn(cout, C, 1, 1, generator=g) / C ** 0.5, 0.1 * torch.randn(cout, generator=g)
    outs = []
    for fused in (True, False):
        pb = PlanBuilder(dtype, cin + C, cout, 1)
        a, res, mid, stats, y = pb.buffer(cin + C), pb.buffer(C), pb.buffer(C), pb.buffer(8), pb.buffer(cout)
        pb.conv(INPUT, a, torch.eye(cin + C).view(cin + C, cin + C, 1, 1))
        pb.conv(a.slice(cin, C), res, torch.eye(C).view(C, C, 1, 1))
        assert pb.ln_out_supported(C)
        pb.conv(a.slice(0, cin), mid, w1, b1, combine=N.COMB_AXPY, res1=res, ln_out=stats if fused else None)
        if not fused:
            pb.layernorm_stats(mid, stats, eps=1e-5)
        pb.conv(mid, y, w2, b2, ln=(stats, gamma, beta) + ((1e-5,) if fused else ()))
        pb.conv(y, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
        got, plan = _run(pb, x, dtype)
        outs.append(got)
        m = plan.read_buffer(mid).double().cpu()  # what the second linear actually read
    ref = F.conv2d(F.layer_norm(m.permute(0, 2, 3, 1), (C,), gamma.double(), beta.double(), 1e-5).permute(0, 3, 1, 2), w2.to(dtype).double(), b2.double())
    _check(outs[0], ref, dtype, what='linear(LayerNorm(.)) with sums from the producing linear')
    _check(outs[1], ref, dtype, what='linear(LayerNorm(.)) with the statistics op')
    span = float(ref.max() - ref.min())
    assert float((outs[0] - outs[1]).abs().max()) <= 1.2e-2 * span


# ------------------------------------------------------------------------------------------------ GateRV3 ops (rsb_op_kind 10-11, sparse depthwise 11x11)
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,B,H,W', [(32, 1, 40, 56), (64, 2, 17, 30), (256, 1, 24, 40), (32, 1, 160, 192)])
def test_channel_gate_and_affine_ops(C, B, H, W, dtype):
    """MetaGated's tail (gaterv3/arch.py:661-665): x * sca(x) * gamma0 + short with sca = Conv1x1(mean_hw(x)), then y * gamma1 + x."""
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, 2 * C, H, W, generator=g) + 0.2
    ws, bs = torch.randn(C, C, 1, 1, generator=g) / C ** 0.5, 0.3 * torch.randn(C, generator=g)
    g0, g1 = 1.0 + 0.3 * torch.randn(1, C, 1, 1, generator=g), 1.0 + 0.3 * torch.randn(1, C, 1, 1, generator=g)
    pb = PlanBuilder(dtype, 2 * C, C, 1)
    a, u, v = pb.buffer(2 * C), pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, a, torch.eye(2 * C).view(2 * C, 2 * C, 1, 1))
    pb.chan_gate(a.slice(0, C), u, a.slice(C, C), ws, bs, g0)
    pb.chan_affine(u, v, g1, res=a.slice(C, C))
    pb.conv(v, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    t, short = xq[:, :C], xq[:, C:]
    sca = F.conv2d(t.mean(dim=(2, 3), keepdim=True), ws.double(), bs.double())
    uu = (t * sca * g0.double() + short).to(dtype).double()
    ref = uu * g1.double() + short
    _check(got, ref, dtype, what='channel gate + affine')


@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,B,H,W', [(32, 1, 40, 56), (48, 2, 9, 30), (512, 1, 12, 20)])
def test_inception_depthwise_11x11(C, B, H, W, dtype):
    """InceptionDWConv2d (gaterv3/arch.py:527-557) as one depthwise 11x11 kernel whose zero taps are skipped per 8-channel plane."""
    from resselt_b200.archs.gaterv3 import merge_inception
    g = torch.Generator().manual_seed(C + W)
    gc = int(C * 0.125)
    x = torch.randn(B, C, H, W, generator=g)
    w = {'t.dwconv_hw.weight': torch.randn(gc, 1, 3, 3, generator=g) / 3, 't.dwconv_hw.bias': 0.1 * torch.randn(gc, generator=g),
         't.dwconv_w.weight': torch.randn(gc, 1, 1, 11, generator=g) / 3, 't.dwconv_w.bias': 0.1 * torch.randn(gc, generator=g),
         't.dwconv_h.weight': torch.randn(gc, 1, 11, 1, generator=g) / 3, 't.dwconv_h.bias': 0.1 * torch.randn(gc, generator=g)}
    pb = PlanBuilder(dtype, C, C, 1)
    a, b = pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.dwconv(a, b, *merge_inception(w, 't', C))
    pb.conv(b, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    D = lambda k: w[k].double()
    c_id, c_hw, c_w, c_h = torch.split(xq, (C - 3 * gc, gc, gc, gc), dim=1)
    ref = torch.cat((c_id, F.conv2d(c_hw, D('t.dwconv_hw.weight'), D('t.dwconv_hw.bias'), padding=1, groups=gc),
                     F.conv2d(c_w, D('t.dwconv_w.weight'), D('t.dwconv_w.bias'), padding=(0, 5), groups=gc),
                     F.conv2d(c_h, D('t.dwconv_h.weight'), D('t.dwconv_h.bias'), padding=(5, 0), groups=gc)), dim=1)
    _check(got, ref, dtype, what='inception depthwise conv')
    if dtype == torch.float32:
        assert torch.equal(got[:, : C - 3 * gc], xq[:, : C - 3 * gc]), 'identity channels pass through bit for bit'


# ------------------------------------------------------------------------------------------------ LayerNorm folded into a linear
@pytest.mark.parametrize('dtype', DTYPES)
@pytest.mark.parametrize('C,cout,act,residual,mean,B,H,W', [
    (180, 180, N.ACT_NONE, False, 0.3, 1, 40, 56),    # norm1 -> q / k / v (swinir/arch.py:268-293)
    (180, 360, N.ACT_GELU, False, 4.0, 2, 33, 47),    # norm2 -> fc1 + GELU; token mean ten times the spread (the term that must cancel)
    (60, 60, N.ACT_NONE, True, -1.0, 1, 64, 64),      # light models; AXPY tail on top of the fold
    (180, 600, N.ACT_NONE, False, 0.5, 1, 24, 40),    # wider than one UMMA N tile: the builder splits the conv, every part carries its row sums
    (180, 180, N.ACT_NONE, False, 0.2, 1, 256, 256),  # many tiles
])
def test_layernorm_folded_into_linear(C, cout, act, residual, mean, B, H, W, dtype):
    """conv(x, ln=(stats, gamma, beta)) == Linear(LayerNorm(x)): statistics pass + epilogue fold (rsb_conv_desc.ln_fold) against fp64."""
    g = torch.Generator().manual_seed(C + cout + H)
    x = torch.randn(B, C, H, W, generator=g) * 0.4 + mean + 0.5 * torch.randn(B, 1, H, W, generator=g)
    gamma, beta = 1.0 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    wl, bl = torch.randn(cout, C, 1, 1, generator=g) / C ** 0.5, 0.1 * torch.randn(cout, generator=g)
    pb = PlanBuilder(dtype, C, cout, 1)
    a, stats, y = pb.buffer(C), pb.buffer(8), pb.buffer(cout)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.layernorm_stats(a, stats, eps=1e-5)
    kw = dict(combine=N.COMB_AXPY, res1=a) if residual else {}
    pb.conv(a, y, wl, bl, act=act, ln=(stats, gamma, beta), **kw)
    pb.conv(y, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    got, _ = _run(pb, x, dtype)
    xq = _q(x, dtype)
    ln = F.layer_norm(xq.permute(0, 2, 3, 1), (C,), gamma.double(), beta.double(), 1e-5).permute(0, 3, 1, 2)
    ref = F.conv2d(ln, wl.double(), bl.double())
    if act == N.ACT_GELU:
        ref = F.gelu(ref)
    if residual:
        ref = ref + xq
    _check(got, ref, dtype, what='LayerNorm fold')
