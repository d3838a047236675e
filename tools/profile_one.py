"""Profiling target: W warm-up + K forwards of one model at one shape (used under ncu).
    python tools/profile_one.py [arch] [h] [w] [warmup] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import DAT, SPAN, RealPLKSR, RRDBNet, SpanPlus, SRVGGNetCompact, SwinIR

arch = sys.argv[1] if len(sys.argv) > 1 else 'span'
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
w = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 1
batch = int(sys.argv[6]) if len(sys.argv) > 6 else 1
dev = torch.device('cuda:0')
model = {'span': lambda: SPAN(feature_channels=48, upscale=2, seed=3),
         'spanplus': lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4),
         'compact': lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5),
         'esrgan': lambda: RRDBNet(num_blocks=23, scale=4, seed=6),
         'plksr': lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=8),
         'dat': lambda: DAT(upscale=4, seed=9),
         'swinir': lambda: SwinIR(upscale=4, seed=10),
         'swinir1': lambda: SwinIR(upscale=4, depths=[2], num_heads=[6], seed=10),
         'spanplus_dys': lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=4),
         'plksr_dys': lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, dysample=True, seed=8),
         'plksr1': lambda: RealPLKSR(n_blocks=1, upscaling_factor=4, seed=8),
         'dat1': lambda: DAT(upscale=4, depth=[4], num_heads=[6], seed=9)}[arch]().eval().to(dev).bfloat16()
x = torch.rand(batch, 3, h, w, device=dev).bfloat16()
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
ms = e0.elapsed_time(e1) / iters
plan = model.plan_for(dev, torch.bfloat16)
print(f'{arch} {batch}x{h}x{w}: {ms:.3f} ms/forward  {batch * h * w * model.upscale ** 2 / ms / 1e3:.1f} out-MP/s  '
      f'{plan.flops(batch, h, w) / ms / 1e9:.1f} TFLOP/s (executed)  launches {plan.launches_per_forward}')
