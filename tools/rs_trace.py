"""One launch of the 48->48 3x3 SiLU layer at 1080p with per-row clock stamps (RSB_RS_TRACE=<file>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.engine import native as N
from resselt_b200.engine.plan import INPUT, OUTPUT, PlanBuilder

pb = PlanBuilder(torch.bfloat16, 3, 3, 1)
a, b = pb.buffer(48), pb.buffer(48)
g = torch.Generator().manual_seed(1)
wt = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
pb.conv(INPUT, a, torch.randn(48, 3, 3, 3, generator=g) * 0.2)
pb.conv(a, b, wt, torch.zeros(48), act=N.ACT_SILU)
pb.conv(b, OUTPUT, torch.randn(3, 48, 3, 3, generator=g) * 0.1)
plan = pb.finalize(torch.device('cuda:0'))
x = torch.rand(1, 3, 1080, 1920, device='cuda:0', dtype=torch.bfloat16)
trace = os.environ.pop('RSB_RS_TRACE', None)
plan.forward(x)
for _ in range(3):
    plan.forward(x, ops=(1, 2))
torch.cuda.synchronize()
if trace:
    os.environ['RSB_RS_TRACE'] = trace
plan.forward(x, ops=(1, 2))
torch.cuda.synchronize()
