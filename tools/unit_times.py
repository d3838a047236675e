"""Per-launch-unit device times of a model (each unit replayed from a CUDA graph), grouped by kernel, next to the whole
forward's graph-replayed time.
    python tools/unit_times.py span|spanplus|compact|esrgan|plksr|dat|swinir [mode] [h w [n]]
mode: 0 = default kernels (fused pairs), 5 = every conv its own launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import DAT, SPAN, GateRV3, RealPLKSR, RRDBNet, RTMoSR, SpanPlus, SpanPP, SRVGGNetCompact, SwinIR
from resselt_b200.engine.profiling import summarize_units, time_forward, time_units

MODELS = {
    'span': (lambda: SPAN(feature_channels=48, upscale=2, seed=3), 1, 1080, 1920),
    'spanplus': (lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4), 1, 1080, 1920),
    'spanplus_dys': (lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=4), 1, 1080, 1920),
    'compact': (lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5), 16, 540, 960),
    'esrgan': (lambda: RRDBNet(num_blocks=23, scale=4, seed=6), 1, 768, 768),
    'plksr': (lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=7), 1, 512, 512),
    'dat': (lambda: DAT(upscale=4, seed=8), 1, 512, 512),
    'swinir': (lambda: SwinIR(upscale=4, seed=9), 1, 512, 512),
    'spanpp': (lambda: SpanPP(feature_channels=48, seed=10), 1, 1080, 1920),
    'rtmosr': (lambda: RTMoSR(scale=2, dim=32, n_blocks=2, seed=11), 1, 1080, 1920),
    'gaterv3': (lambda: GateRV3(scale=2, seed=12), 1, 1088, 1920),
}

if __name__ == '__main__':
    name = sys.argv[1] if len(sys.argv) > 1 else 'span'
    mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    make, n, h, w = MODELS[name]
    if len(sys.argv) > 4:
        h, w = int(sys.argv[3]), int(sys.argv[4])
    if len(sys.argv) > 5:
        n = int(sys.argv[5])
    dev = torch.device('cuda:0')
    m = make().eval().to(dev).bfloat16()
    x = torch.rand(n, 3, h, w, device=dev).bfloat16()
    plan = m.plan_for(dev, torch.bfloat16)
    plan.force_direct = mode
    with torch.inference_mode():
        out = torch.empty((n, m.out_channels, h * m.upscale, w * m.upscale), dtype=torch.bfloat16, device=dev)
        fwd = time_forward(plan, x, out, reps=5)
        units = time_units(plan, x, out, reps=10)
    for u in units:
        print(f"ops {u['ops'][0]:3d}-{u['ops'][1]:3d} {u['kernel']:12s} {u['ms'] * 1e3:8.1f} us  {u['flops'] / max(u['ms'], 1e-9) / 1e9:8.1f} TFLOP/s  {u['bytes'] / max(u['ms'], 1e-9) / 1e6:8.1f} GB/s")
    print(json.dumps(dict(model=name, mode=mode, shape=[n, 3, h, w], forward_ms=fwd, out_mp_per_s=n * h * w * m.upscale ** 2 / 1e6 / (fwd * 1e-3),
                          launches=plan.launches_per_forward, fused_pairs=plan.fused_pairs, kernels=summarize_units(units, fwd))))
