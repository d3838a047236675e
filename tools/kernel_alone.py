"""Dominant kernel of the headline bench (SPAN block_1.c1_r: 3x3 48->48 + SiLU at 1080p) timed alone two ways:
Python launch loop vs one CUDA-graph replay of the same launches (no host in the loop)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import SPAN

dev = torch.device('cuda:0')
m = SPAN(feature_channels=48, upscale=2, seed=3).eval().to(dev).bfloat16()
x = torch.rand(1, 3, 1080, 1920, device=dev).bfloat16()
plan = m.plan_for(dev, torch.bfloat16)
out = torch.empty(1, 3, 2160, 3840, device=dev, dtype=torch.bfloat16)
iters = 50
with torch.inference_mode():
    for _ in range(3):
        plan.forward(x, out=out)
    torch.cuda.synchronize()
    for op in (1, 2, 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            plan.forward(x, out=out, ops=(op, op + 1))
        e1.record()
        torch.cuda.synchronize()
        loop_us = e0.elapsed_time(e1) / iters * 1e3
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                plan.forward(x, out=out, ops=(op, op + 1))
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f'op {op}: python loop {loop_us:.1f} us/launch, graph replay {e0.elapsed_time(e1) / iters * 1e3:.1f} us/launch')
    # whole forward as a graph
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            plan.forward(x, out=out)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'whole forward, graph of 10: {e0.elapsed_time(e1) / 10:.3f} ms')
    e0.record()
    for _ in range(10):
        plan.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f'whole forward, python loop of 10: {e0.elapsed_time(e1) / 10:.3f} ms')
