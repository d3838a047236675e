"""GPU tests of the fused conv-pair kernel (csrc/conv_pair.cu, opt-in plan mode 4): two consecutive 3x3 convs in one launch,
the intermediate map kept in shared memory.  The fused pair must equal the two separate row-streaming launches BIT FOR BIT
(same accumulation order per pixel, same bf16 rounding of the intermediate).
Plan modes (include/resselt_b200.h): 0 = default, 3 = row-streaming kernel per conv, 4 = default + fused pairs."""
import pytest
import torch

from resselt_b200.archs import SPAN, SpanPlus
from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _pair_plan(act, gate, seed=0):
    g = torch.Generator().manual_seed(seed)
    pb = PlanBuilder(torch.bfloat16, 48, 48, 1)
    a, b, c = pb.buffer(48), pb.buffer(48), pb.buffer(48)
    w1 = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
    w2 = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
    pb.conv(INPUT, a, torch.eye(48).view(48, 48, 1, 1))
    pb.conv(a, b, w1, torch.randn(48, generator=g) * 0.1, act=act)
    if gate:
        pb.conv(b, c, w2, torch.randn(48, generator=g) * 0.1, combine=N.COMB_SPAB_GATE, res1=a)
    else:
        pb.conv(b, c, w2, torch.randn(48, generator=g) * 0.1, act=act)
    pb.conv(c, OUTPUT, torch.eye(48).view(48, 48, 1, 1))
    return pb.finalize(torch.device(DEV)), c


@pytest.mark.parametrize('n,H,W,act,gate', [
    (1, 300, 512, N.ACT_SILU, True),    # five strips of 120 owned pixels, CTA runs start and end inside strips
    (1, 1200, 128, N.ACT_SILU, True),   # second strip owns 7 pixels
    (2, 400, 248, N.ACT_MISH, True),    # two images
    (1, 1500, 120, N.ACT_NONE, False),  # a single strip narrower than the 128-pixel MMA segment
    (3, 211, 256, N.ACT_SILU, False),   # image height not a multiple of anything
])
def test_fused_pair_equals_two_launches(n, H, W, act, gate):
    plan, c = _pair_plan(act, gate)
    x = torch.randn(n, 48, H, W, generator=torch.Generator().manual_seed(H * 7 + W)).to(DEV, torch.bfloat16)
    got = {}
    for mode, flag in (('pair', 4), ('rs', 3)):
        plan.force_direct = flag
        plan.forward(x)
        torch.cuda.synchronize()
        got[mode] = plan.read_buffer(c).cpu()
    plan.force_direct = 0
    assert torch.isfinite(got['rs']).all() and float(got['rs'].abs().max()) > 0.1
    assert torch.equal(got['pair'], got['rs']), f'max |pair - rs| = {float((got["pair"] - got["rs"]).abs().max()):.3e}'


def _chain_plan(act, seed=0):
    """stem(16->48, output kept) -> [c1 act, c2 act, c3 gate(x)] x 2 -> tail conv: every pair flavour of a SPAN chain —
    (stem+store, act), (act, gate), (gate+store, act), (act+store... no: act, act), (gate, none)."""
    g = torch.Generator().manual_seed(seed)

    def w(cin):
        return torch.randn(48, cin, 3, 3, generator=g) / (cin * 9) ** 0.5, torch.randn(48, generator=g) * 0.1

    pb = PlanBuilder(torch.bfloat16, 16, 48, 1)
    x0, f, t1, t2, b1, b2, o = (pb.buffer(c) for c in (16, 48, 48, 48, 48, 48, 48))
    pb.conv(INPUT, x0, torch.eye(16).view(16, 16, 1, 1))
    pb.conv(x0, f, *w(16))                                            # stem: read again by the first gate -> stored
    pb.conv(f, t1, *w(48), act=act)
    pb.conv(t1, t2, *w(48), act=act)
    pb.conv(t2, b1, *w(48), combine=N.COMB_SPAB_GATE, res1=f)         # block output: next gate's residual -> stored
    pb.conv(b1, t1, *w(48), act=act)                                  # t1 is read by the final 1x1 as well -> stored
    pb.conv(t1, t2, *w(48), act=act)
    pb.conv(t2, b2, *w(48), combine=N.COMB_SPAB_GATE, res1=b1)
    pb.conv(b2, o, *w(48))
    cat = pb.buffer(96)
    pb.conv(o, cat.slice(0, 48), torch.eye(48).view(48, 48, 1, 1))
    pb.conv(t1, cat.slice(48, 48), torch.eye(48).view(48, 48, 1, 1))
    pb.conv(cat, OUTPUT, torch.randn(48, 96, 1, 1, generator=g) / 96 ** 0.5)
    return pb.finalize(torch.device(DEV)), (f, b1, t1, o)


@pytest.mark.parametrize('n,H,W,act', [
    (1, 640, 512, N.ACT_SILU),
    (2, 333, 376, N.ACT_MISH),
    (1, 1300, 120, N.ACT_SILU),
])
def test_chain_of_pairs_equals_single_launches(n, H, W, act):
    plan, keep = _chain_plan(act)
    x = torch.randn(n, 16, H, W, generator=torch.Generator().manual_seed(H + W)).to(DEV, torch.bfloat16)
    got = {}
    for mode in (4, 0):
        plan.force_direct = mode
        y = plan.forward(x).float().cpu()
        torch.cuda.synchronize()
        got[mode] = [y] + [plan.read_buffer(r).cpu() for r in keep]
        if mode == 4:
            fused = plan.fused_pairs
    assert plan.num_ops == 12 and fused == 4, f'expected 4 fused pairs, got {fused}'
    for a, b in zip(got[4], got[0]):
        assert torch.isfinite(b).all() and float(b.abs().max()) > 0.05
        assert torch.equal(a, b), f'max |pair - single| = {float((a - b).abs().max()):.3e}'
    plan.force_direct = 0


@pytest.mark.parametrize('model,shape,pairs', [
    (SPAN(feature_channels=48, upscale=2, seed=41), (1, 3, 600, 512), 10),   # conv_1 .. conv_2: 20 convs = 10 pairs
    (SpanPlus(blocks=[2], feature_channels=48, upscale=2, seed=42), (2, 3, 360, 480), 7),
    (SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=43), (1, 3, 400, 640), 10),
])
def test_whole_model_with_fused_pairs_is_bit_identical(model, shape, pairs):
    m = model.eval().to(DEV).bfloat16()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(5)).to(DEV, torch.bfloat16)
    plan = m.plan_for(torch.device(DEV), torch.bfloat16)
    with torch.inference_mode():
        plan.force_direct = 0
        y_ref = m(x).clone()
        assert plan.fused_pairs == 0
        plan.force_direct = 4
        y_pair = m(x).clone()
        fused = plan.fused_pairs
        plan.force_direct = 0
    assert fused == pairs, f'{fused} fused pairs, expected {pairs}'
    assert torch.equal(y_pair, y_ref)
