"""GPU tests of the row-streaming 3x3 kernel (csrc/conv_rs.cu): every eligible conv must equal the tile kernel
(csrc/conv_tc.cu) BIT FOR BIT — the two formulations accumulate each output pixel in the same order, which is what keeps
tiled and untiled forwards identical when tiles fall on different kernels (conv_rs needs W % 8 == 0) — and match an
fp64 reference of the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from resselt_b200.archs import SPAN, SpanPlus, SRVGGNetCompact
from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

CASES = [
    # cin, cout, n, H, W, act, combine
    (48, 48, 1, 32, 40, N.ACT_NONE, N.COMB_NONE),     # one column strip narrower than the 144-pixel TMA box
    (48, 48, 1, 1, 8, N.ACT_NONE, N.COMB_NONE),       # a single row: first == last image row
    (48, 48, 1, 2, 136, N.ACT_SILU, N.COMB_NONE),     # second column strip is 8 pixels wide
    (48, 48, 2, 300, 256, N.ACT_SILU, N.COMB_NONE),   # CTA ranges start and end inside columns; several ring wraps
    (48, 48, 1, 64, 384, N.ACT_NONE, N.COMB_SPAB_GATE),
    (48, 48, 3, 7, 128, N.ACT_MISH, N.COMB_NONE),     # columns shorter than the accumulator ring
    (48, 48, 1, 40, 64, N.ACT_NONE, N.COMB_MUL),
    (64, 64, 2, 33, 24, N.ACT_LRELU, N.COMB_AXPY),
    (64, 64, 1, 50, 72, N.ACT_PRELU, N.COMB_NONE),
    (160, 32, 1, 40, 160, N.ACT_LRELU, N.COMB_NONE),  # ESRGAN dense-block shape
    (48, 12, 1, 32, 24, N.ACT_NONE, N.COMB_NONE),     # narrow N (16): 32-slot ring
    (80, 40, 1, 19, 32, N.ACT_SIGMOID, N.COMB_NONE),  # channel counts that are not multiples of 16
    (16, 80, 1, 23, 16, N.ACT_GELU, N.COMB_NONE),     # widest eligible N (3 x 80 = 240): 6-slot ring
]


@pytest.mark.parametrize('cin,cout,n,H,W,act,comb', CASES)
def test_row_streaming_equals_tile_kernel_and_fp64(cin, cout, n, H, W, act, comb):
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    x = torch.randn(n, cin, H, W, generator=g)
    res = torch.randn(n, cout, H, W, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    slopes = torch.rand(cout, generator=g) * 0.5
    pb = PlanBuilder(torch.bfloat16, cin + cout, cout, 1)
    a, r, b = pb.buffer(cin), pb.buffer(cout), pb.buffer(cout)
    eye = torch.eye(cin + cout)
    pb.conv(INPUT, a, eye[:cin].reshape(cin, cin + cout, 1, 1))
    pb.conv(INPUT, r, eye[cin:].reshape(cout, cin + cout, 1, 1))
    kw = dict(act=act, act_param=0.2, combine=comb)
    if act == N.ACT_PRELU:
        kw['act_slopes'] = slopes
    if comb != N.COMB_NONE:
        kw.update(res1=r, alpha=0.2, beta1=1.0)
    pb.conv(a, b, wt, bias, **kw)
    pb.conv(b, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    plan = pb.finalize(torch.device(DEV))
    xd = torch.cat([x, res], dim=1).to(DEV, torch.bfloat16)
    got = {}
    for mode, flag in (('rs', 3), ('tc', 2)):
        plan.force_direct = flag
        plan.forward(xd)
        got[mode] = plan.read_buffer(b).cpu()
    plan.force_direct = 0
    assert torch.equal(got['rs'], got['tc']), f'max |rs - tc| = {float((got["rs"] - got["tc"]).abs().max()):.3e}'
    q = lambda t: t.to(torch.bfloat16).double()
    v = F.conv2d(q(x), q(wt), bias.double(), padding=1)
    if comb == N.COMB_SPAB_GATE:
        ref = (v + q(res)) * (torch.sigmoid(v) - 0.5)
    else:
        ref = {N.ACT_NONE: lambda t: t, N.ACT_SILU: F.silu, N.ACT_MISH: F.mish, N.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2),
               N.ACT_GELU: F.gelu, N.ACT_SIGMOID: torch.sigmoid,
               N.ACT_PRELU: lambda t: torch.where(t >= 0, t, t * slopes.double().view(1, -1, 1, 1))}[act](v)
        if comb == N.COMB_MUL:
            ref = ref * q(res)
        elif comb == N.COMB_AXPY:
            ref = 0.2 * ref + q(res)
    assert float((got['rs'].double() - ref).abs().max()) / float(ref.abs().max()) < 8e-3  # bf16 output rounding + fp32 accumulation


@pytest.mark.parametrize('model,shape', [
    (SPAN(feature_channels=48, upscale=2, seed=41), (1, 3, 120, 200)),
    (SpanPlus(blocks=[2], feature_channels=48, upscale=2, seed=42), (2, 3, 64, 72)),
    (SRVGGNetCompact(num_feat=64, num_conv=4, upscale=4, seed=43), (1, 3, 56, 136)),
])
def test_whole_model_is_identical_on_both_tensor_core_kernels(model, shape):
    m = model.eval().to(DEV).bfloat16()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(5)).to(DEV, torch.bfloat16)
    plan = m.plan_for(torch.device(DEV), torch.bfloat16)
    with torch.inference_mode():
        plan.force_direct = 3
        y_rs = m(x).clone()
        plan.force_direct = 2
        y_tc = m(x).clone()
        plan.force_direct = 0
    assert torch.equal(y_rs, y_tc)
