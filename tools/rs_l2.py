"""Does a ping-pong chain of 48->48 3x3 layers stay in L2 when the maps are small enough?
Times a->b, b->a alternately for several image heights (width 1920): per-layer time vs working set."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.engine import native as N
from resselt_b200.engine.plan import INPUT, OUTPUT, PlanBuilder

DEV = 'cuda:0'
for H in (1080, 540, 360, 270, 216, 180, 135, 90):
    pb = PlanBuilder(torch.bfloat16, 3, 3, 1)
    a, b = pb.buffer(48), pb.buffer(48)
    g = torch.Generator().manual_seed(1)
    wt = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
    pb.conv(INPUT, a, torch.randn(48, 3, 3, 3, generator=g) * 0.2)
    pb.conv(a, b, wt, torch.zeros(48), act=N.ACT_SILU)
    pb.conv(b, a, wt, torch.zeros(48), act=N.ACT_SILU)
    pb.conv(b, OUTPUT, torch.randn(3, 48, 3, 3, generator=g) * 0.1)
    plan = pb.finalize(torch.device(DEV))
    x = torch.rand(1, 3, H, 1920, device=DEV, dtype=torch.bfloat16)
    plan.forward(x)
    for _ in range(3):
        plan.forward(x, ops=(1, 3))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        plan.forward(x, ops=(1, 3))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (2 * reps)
    mb = H * 1920 * 48 * 2 / 1e6
    fl = 2 * 48 * 48 * 9 * H * 1920
    print(f'H={H:5d}: map {mb:6.1f} MB  {us:7.1f} us/layer  {fl / us / 1e6:7.1f} TFLOP/s  ({us * 1080 / H:6.1f} us per 1080p-equivalent)', flush=True)
