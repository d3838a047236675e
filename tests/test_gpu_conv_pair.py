"""GPU tests of the experimental fused conv-pair kernel (csrc/conv_pair.cu, plan mode 4 / RSB_PAIR=1): two consecutive
3x3 convs in one launch, the intermediate map kept in shared memory.  The fused pair must equal the two separate
row-streaming launches BIT FOR BIT (same accumulation order per pixel, same bf16 rounding of the intermediate)."""
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


@pytest.mark.parametrize('model,shape', [
    (SPAN(feature_channels=48, upscale=2, seed=41), (1, 3, 600, 512)),      # six c2_r -> c3_r + gate pairs
    (SpanPlus(blocks=[2], feature_channels=48, upscale=2, seed=42), (2, 3, 360, 480)),
])
def test_whole_model_with_fused_pairs_is_bit_identical(model, shape):
    m = model.eval().to(DEV).bfloat16()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(5)).to(DEV, torch.bfloat16)
    plan = m.plan_for(torch.device(DEV), torch.bfloat16)
    with torch.inference_mode():
        plan.force_direct = 4
        y_pair = m(x).clone()
        plan.force_direct = 0
        y_ref = m(x).clone()
    assert torch.equal(y_pair, y_ref)
