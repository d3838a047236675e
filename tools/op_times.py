"""Per-op device time of one model forward, ops launched in plan order (so each op sees the cache state the previous
one left) with a CUDA event pair around every op.   python tools/op_times.py [arch] [h] [w] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import DAT, SPAN, RealPLKSR, RRDBNet, SpanPlus, SRVGGNetCompact, SwinIR

arch = sys.argv[1] if len(sys.argv) > 1 else 'span'
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device('cuda:0')
m = {'span': lambda: SPAN(feature_channels=48, upscale=2, seed=3),
     'spanplus': lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4),
     'compact': lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5),
     'esrgan': lambda: RRDBNet(num_blocks=23, scale=4, seed=6),
     'spanplus_dys': lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=4),
     'plksr': lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=7),
     'dat': lambda: DAT(upscale=4, seed=8),
     'swinir': lambda: SwinIR(upscale=4, seed=9)}[arch]().eval().to(dev).bfloat16()
x = torch.rand(1, 3, h, w, device=dev).bfloat16()
plan = m.plan_for(dev, torch.bfloat16)
out = torch.empty(1, 3, m.upscale * h, m.upscale * w, device=dev, dtype=torch.bfloat16)
n = plan.num_ops
for _ in range(2):
    plan.forward(x, out=out)
torch.cuda.synchronize()
tot = [0.0] * n
for r in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for op in range(n):
        plan.forward(x, out=out, ops=(op, op + 1))
        ev[op + 1].record()
    torch.cuda.synchronize()
    for op in range(n):
        tot[op] += ev[op].elapsed_time(ev[op + 1]) * 1e3 / reps
for op in range(n):
    print(f'op {op:3d}: {tot[op]:8.1f} us')
print(f'sum {sum(tot) / 1e3:.3f} ms over {n} ops (includes host launch gaps inside each event pair)')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.forward(x, out=out)
e1.record()
torch.cuda.synchronize()
print(f'whole forward: {e0.elapsed_time(e1) / reps:.3f} ms')
