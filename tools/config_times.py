"""Device-resident forward time of every BASELINE.json configuration on one B200 (bf16), one JSON line per config.
Untiled forwards are replayed from a CUDA graph (no host launch cost: DAT is 675 launches); the tiled case runs eagerly.
    python tools/config_times.py [iters] [case,case,...]
Tensor-core ceilings: algorithmic FLOP per output pixel x out-MP/s against the sustained bf16 peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.archs import DAT, SPAN, GateRV3, RealPLKSR, RRDBNet, RTMoSR, SpanPlus, SpanPP, SRVGGNetCompact, SwinIR
from resselt_b200.engine.profiling import time_forward
from resselt_b200.runner import tiled_forward

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else None  # optional: indices of the cases to run
dev = torch.device('cuda:0')
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
SUSTAINED = peaks.get('bf16_tflops_sustained', 1385.7) * 1e12
# (label, model factory, batch, h, w, algorithmic FLOP per INPUT pixel, tile or None)
CASES = [
    ('config1/3 SPAN 2x 1080p', lambda: SPAN(feature_channels=48, upscale=2, seed=3), 1, 1080, 1920, 819360, None),
    ('config3 SPANPlus 2x [4] 1080p', lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4), 1, 1080, 1920, 819360, None),
    ('config2 Compact 4x nf64 nc16, 16 x 540p', lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5), 16, 540, 960, 1238400, None),
    ('config2 Compact 4x, 1 x 540p', lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5), 1, 540, 960, 1238400, None),
    ('config4 ESRGAN 4x nb23, one 768x768 tile', lambda: RRDBNet(num_blocks=23, scale=4, seed=6), 1, 768, 768, 35853696, None),
    ('config4 ESRGAN 4x nb23, 1080x1920 image as 2x2 exact-halo tiles (halo 349)', lambda: RRDBNet(num_blocks=23, scale=4, seed=6), 1, 1080, 1920, 35853696, (540, 960)),
    ('config5 RealPLKSR 4x 512^2', lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=7), 1, 512, 512, 14753152, None),
    ('config5 DAT 4x 512^2', lambda: DAT(upscale=4, seed=8), 1, 512, 512, 26242128, None),
    ('8a-a19 SwinIR 4x (180ch 6x6 w8) 512^2', lambda: SwinIR(upscale=4, seed=9), 1, 512, 512, None, None),
    ('8f-2 SpanPP 2x 1080p', lambda: SpanPP(feature_channels=48, seed=10), 1, 1080, 1920, None, None),
    ('8f-2 RTMoSR 2x (dim 32, 2 blocks) 1080p', lambda: RTMoSR(scale=2, dim=32, n_blocks=2, seed=11), 1, 1080, 1920, None, None),
    ('8f-2 GateRV3 2x (dim 32, (2,2,4,8) U-Net, 12 latent blocks) 1088x1920 (1080p reflect-padded to 16)', lambda: GateRV3(scale=2, seed=12), 1, 1088, 1920, None, None),
    ('8f-4 SPANPlus 2x dys head 1080p', lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=4), 1, 1080, 1920, None, None),
]
for idx, (label, make, b, h, w, flop_px, tile) in enumerate(CASES):
    if only is not None and idx not in only:
        continue
    m = make().eval().to(dev).bfloat16()
    x = torch.rand(b, 3, h, w, device=dev).bfloat16()
    with torch.inference_mode():
        if tile:
            run = lambda: tiled_forward(m, x, m.upscale, tile, m.receptive_radius)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
        else:
            out = torch.empty((b, m.out_channels, h * m.upscale, w * m.upscale), dtype=torch.bfloat16, device=dev)
            ms = time_forward(m.plan_for(dev, torch.bfloat16), x, out, reps=iters)
    out_mp = b * h * w * m.upscale ** 2 / 1e6
    rec = dict(config=label, ms=round(ms, 3), out_mp_per_s=round(out_mp / ms * 1e3, 1))
    if flop_px:
        rec['algorithmic_tflops'] = round(flop_px * b * h * w / ms / 1e9, 1)
        rec['frac_of_sustained_bf16'] = round(flop_px * b * h * w / (ms * 1e-3) / SUSTAINED, 3)
    print(json.dumps(rec), flush=True)
    del m, x
    torch.cuda.empty_cache()
