"""GPU tests of the large-kernel row-streaming conv (csrc/conv_lk.cu; RealPLKSR's dense 17x17 conv,
/root/reference/resselt/archs/plksr/rplksr.py:27,36): it must equal the tap-by-tap tile kernel (csrc/conv_tc.cu) BIT FOR
BIT — both accumulate each output pixel in (kh, kw, k) order — and match an fp64 reference of the bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from resselt_b200.archs import RealPLKSR
from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('k,cin,cout,n,H,W,act', [
    (17, 16, 16, 1, 64, 128, N.ACT_NONE),    # RealPLKSR's shape; one strip
    (17, 16, 16, 2, 300, 264, N.ACT_NONE),   # three strips (the last 8 pixels wide), runs starting inside strips, ring wraps
    (17, 16, 16, 1, 5, 40, N.ACT_NONE),      # image lower than the kernel: every row sees both borders
    (13, 16, 16, 1, 90, 136, N.ACT_SILU),    # other odd extents
    (7, 32, 16, 1, 70, 72, N.ACT_NONE),      # two K steps
    (5, 16, 8, 3, 33, 24, N.ACT_LRELU),      # half-filled N
])
def test_large_kernel_equals_tile_kernel_and_fp64(k, cin, cout, n, H, W, act):
    g = torch.Generator().manual_seed(k * 100 + cin + H)
    x = torch.randn(n, cin, H, W, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    pb = PlanBuilder(torch.bfloat16, cin, cout, 1)
    a, b = pb.buffer(cin), pb.buffer(cout)
    pb.conv(INPUT, a, torch.eye(cin).view(cin, cin, 1, 1))
    pb.conv(a, b, wt, bias, act=act, act_param=0.2)
    pb.conv(b, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    plan = pb.finalize(torch.device(DEV))
    xd = x.to(DEV, torch.bfloat16)
    got = {}
    for mode, flag in (('lk', 3), ('tc', 2)):
        plan.force_direct = flag
        plan.forward(xd)
        torch.cuda.synchronize()
        got[mode] = plan.read_buffer(b).cpu()
    plan.force_direct = 0
    assert torch.equal(got['lk'], got['tc']), f'max |lk - tc| = {float((got["lk"] - got["tc"]).abs().max()):.3e}'
    q = lambda t: t.to(torch.bfloat16).double()
    ref = F.conv2d(q(x), q(wt), bias.double(), padding=k // 2)
    ref = {N.ACT_NONE: lambda t: t, N.ACT_SILU: F.silu, N.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2)}[act](ref)
    assert float((got['lk'].double() - ref).abs().max()) / float(ref.abs().max()) < 8e-3


def test_realplksr_is_identical_with_and_without_the_large_kernel_path():
    m = RealPLKSR(n_blocks=2, upscaling_factor=2, seed=44).eval().to(DEV).bfloat16()
    x = torch.rand(1, 3, 96, 136, generator=torch.Generator().manual_seed(6)).to(DEV, torch.bfloat16)
    plan = m.plan_for(torch.device(DEV), torch.bfloat16)
    with torch.inference_mode():
        plan.force_direct = 3
        y_lk = m(x).clone()
        plan.force_direct = 2
        y_tc = m(x).clone()
        plan.force_direct = 0
    assert torch.equal(y_lk, y_tc)
