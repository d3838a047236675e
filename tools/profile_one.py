"""Profiling target: W warm-up + K forwards of one model at one shape (used under ncu).
    python tools/profile_one.py [arch] [h] [w] [warmup] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import SPAN, SpanPlus, SRVGGNetCompact

arch = sys.argv[1] if len(sys.argv) > 1 else 'span'
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device('cuda:0')
model = {'span': lambda: SPAN(feature_channels=48, upscale=2, seed=3),
         'spanplus': lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4),
         'compact': lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5)}[arch]().eval().to(dev).bfloat16()
x = torch.rand(1, 3, h, w, device=dev).bfloat16()
with torch.inference_mode():
    for _ in range(warm):
        model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = model(x)
    e1.record()
    torch.cuda.synchronize()
print(f'{arch} {h}x{w}: {e0.elapsed_time(e1) / iters:.3f} ms/forward')
