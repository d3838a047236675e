"""The reference on the same GPU: eager PyTorch bf16 (cuDNN / cuBLAS), NCHW and channels_last, next to the engine, for every
BASELINE.json configuration ("the GPU baseline to beat on the same box").  Needs the reference installed
in baseline/_ref (tools/install_reference.sh).  One JSON line per configuration.
    python tools/library_baseline.py [case,case,...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
import torch

import resselt  # the unmodified reference
from resselt_b200.archs import DAT, SPAN, GateRV3, RealPLKSR, RRDBNet, RTMoSR, SpanPlus, SRVGGNetCompact, SwinIR
from resselt_b200.engine.profiling import time_forward

dev = torch.device('cuda:0')
CASES = [
    ('SPAN 2x 1080p', lambda: SPAN(feature_channels=48, upscale=2, seed=3), 1, 1080, 1920),
    ('SPANPlus 2x [4] 1080p', lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4), 1, 1080, 1920),
    ('Compact 4x nf64 nc16, 16 x 540p', lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5), 16, 540, 960),
    ('ESRGAN 4x nb23, 768x768 tile', lambda: RRDBNet(num_blocks=23, scale=4, seed=6), 1, 768, 768),
    ('RealPLKSR 4x 512^2', lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=7), 1, 512, 512),
    ('DAT 4x 512^2', lambda: DAT(upscale=4, seed=8), 1, 512, 512),
    ('SwinIR 4x 512^2', lambda: SwinIR(upscale=4, seed=9), 1, 512, 512),
    ('RTMoSR 2x (dim 32, 2 blocks) 1080p', lambda: RTMoSR(scale=2, dim=32, n_blocks=2, seed=11), 1, 1080, 1920),
    ('GateRV3 2x (dim 32) 1088x1920', lambda: GateRV3(scale=2, seed=12), 1, 1088, 1920),
]
only = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else None
torch.backends.cudnn.benchmark = True
for idx, (label, make, b, h, w) in enumerate(CASES):
    if only is not None and idx not in only:
        continue
    proto = make()
    sd = {k: v.clone() for k, v in proto.state_dict().items()}
    x = torch.rand(b, 3, h, w, device=dev).bfloat16()
    rec = dict(config=label)
    with torch.inference_mode():
        eng = proto.eval().to(dev).bfloat16()
        out = torch.empty((b, eng.out_channels, h * eng.upscale, w * eng.upscale), dtype=torch.bfloat16, device=dev)
        rec['engine_ms'] = round(time_forward(eng.plan_for(dev, torch.bfloat16), x, out, reps=5), 3)
        del eng, out
        torch.cuda.empty_cache()
        for name, fmt in (('eager_nchw_ms', torch.contiguous_format), ('eager_channels_last_ms', torch.channels_last)):
            try:
                ref = resselt.load_from_state_dict({k: v.clone() for k, v in sd.items()}).eval().to(dev).bfloat16().to(memory_format=fmt)
                xx = x.to(memory_format=fmt)
                for _ in range(2):
                    ref(xx)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    ref(xx)
                e1.record()
                torch.cuda.synchronize()
                rec[name] = round(e0.elapsed_time(e1) / 3, 3)
                del ref
            except Exception as exc:  # noqa: BLE001
                rec[name] = f'{type(exc).__name__}: {str(exc)[:120]}'
            torch.cuda.empty_cache()
        if not any(isinstance(rec.get(k), float) for k in ('eager_nchw_ms', 'eager_channels_last_ms')):
            # the reference module itself cannot run with bf16 parameters (DAT / SwinIR build their masks and index tensors in
            # fp32): the library baselines are then bf16 autocast over fp32 parameters, and plain fp32 (TF32 convs, cuDNN default)
            for name, ctx in (('eager_autocast_bf16_ms', lambda: torch.autocast('cuda', dtype=torch.bfloat16)), ('eager_fp32_ms', None)):
                try:
                    ref = resselt.load_from_state_dict({k: v.clone() for k, v in sd.items()}).eval().to(dev)
                    xx = x.float()

                    def run():
                        if ctx is None:
                            return ref(xx)
                        with ctx():
                            return ref(xx)

                    for _ in range(2):
                        run()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    rec[name] = round(e0.elapsed_time(e1) / 3, 3)
                    del ref
                except Exception as exc:  # noqa: BLE001
                    rec[name] = f'{type(exc).__name__}: {str(exc)[:120]}'
                torch.cuda.empty_cache()
    times = [v for k, v in rec.items() if k.startswith('eager') and isinstance(v, float)]
    if times:
        rec['engine_speedup_over_best_eager'] = round(min(times) / rec['engine_ms'], 2)
    print(json.dumps(rec), flush=True)
