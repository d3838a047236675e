"""Bring-up diagnostics for a GPU box (not part of the test suite): layer-level checks of the tensor-core
conv against the CUDA-core conv and a CPU fp64 reference, whole-model parity, and a timing probe.

    python tools/gpu_check.py layers | models | time | all
Each stage runs in its own subprocess with a timeout so a hang or fault in one cannot take the others down.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bf16(t):
    import torch
    return t.to(torch.bfloat16).to(torch.float64)


def stage_layers():
    import torch
    import torch.nn.functional as F
    from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
    from resselt_b200.engine import native as N

    dev = torch.device('cuda:0')
    cases = [
        # cin, cout, k, n, H, W, act
        (48, 48, 3, 1, 32, 40, N.ACT_NONE),
        (48, 48, 3, 1, 37, 45, N.ACT_SILU),
        (64, 64, 3, 2, 33, 17, N.ACT_MISH),
        (48, 12, 3, 1, 32, 24, N.ACT_NONE),
        (192, 48, 1, 1, 40, 40, N.ACT_NONE),
        (16, 16, 17, 1, 48, 40, N.ACT_NONE),
        (64, 128, 3, 1, 32, 32, N.ACT_LRELU),
        (32, 256, 3, 1, 16, 24, N.ACT_NONE),
    ]
    ok_all = True
    for (cin, cout, k, n, H, W, act) in cases:
        g = torch.Generator().manual_seed(cin * 1000 + cout + k)
        x = torch.randn(n, cin, H, W, generator=g)
        wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        bias = torch.randn(cout, generator=g)
        pb = PlanBuilder(torch.bfloat16, cin, cout, 1)
        a, b = pb.buffer(cin), pb.buffer(cout)
        eye_in = torch.eye(cin).view(cin, cin, 1, 1)
        eye_out = torch.eye(cout).view(cout, cout, 1, 1)
        pb.conv(INPUT, a, eye_in)
        pb.conv(a, b, wt, bias, act=act, act_param=0.2)
        pb.conv(b, OUTPUT, eye_out)
        plan = pb.finalize(dev)
        xd = x.to(dev, torch.bfloat16)
        res = {}
        for mode in ('tc', 'direct'):
            plan.force_direct = mode == 'direct'
            y = plan.forward(xd)
            torch.cuda.synchronize()
            res[mode] = (plan.read_buffer(b).double().cpu(), y.double().cpu())
        ref = F.conv2d(_bf16(x), _bf16(wt), bias.double(), padding=k // 2)
        if act == N.ACT_SILU:
            ref = F.silu(ref)
        elif act == N.ACT_MISH:
            ref = F.mish(ref)
        elif act == N.ACT_LRELU:
            ref = F.leaky_relu(ref, 0.2)
        scale = ref.abs().max().item()
        e_tc = (res['tc'][0] - ref).abs().max().item() / scale
        e_dir = (res['direct'][0] - ref).abs().max().item() / scale
        e_out = (res['tc'][1] - res['tc'][0]).abs().max().item() / scale
        ok = e_tc < 2e-2 and e_dir < 2e-2 and e_out < 1e-2
        ok_all &= ok
        print(json.dumps(dict(stage='layer', cin=cin, cout=cout, k=k, n=n, H=H, W=W, act=act, err_tc=e_tc, err_direct=e_dir,
                              err_out_vs_buf=e_out, ok=ok)), flush=True)
        if not ok:
            d = (res['tc'][0] - ref).abs()
            idx = torch.nonzero(d > 2e-2 * scale)
            print('  first mismatches (n,c,y,x):', idx[:8].tolist(), 'count', idx.shape[0], 'of', d.numel(), flush=True)
    return ok_all


def stage_models():
    import torch
    import oracle
    import resselt_b200
    from resselt_b200.archs import DAT, SPAN, SpanPlus, SRVGGNetCompact, RRDBNet, RealPLKSR

    dev = 'cuda:0'
    ok_all = True
    torch.manual_seed(1)
    x = torch.rand(1, 3, 64, 96)
    only = os.environ.get('RSB_CHECK_ONLY')
    models = [
        ('SPAN', SPAN(feature_channels=48, upscale=2, seed=3)),
        ('SPANPlus', SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4)),
        ('Compact', SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5)),
        ('ESRGAN', RRDBNet(num_blocks=3, scale=4, seed=6)),
        ('ESRGAN', RRDBNet(num_blocks=1, scale=2, plus=True, seed=7)),
        ('RealPLKSR', RealPLKSR(n_blocks=3, upscaling_factor=4, seed=8)),
        ('DAT', DAT(depth=[3, 3], num_heads=[6, 6], upscale=4, seed=9)),
    ]
    for name, proto in models:
        if only and name != only:
            continue
        sd = {k: v.clone() for k, v in proto.state_dict().items()}
        ref = oracle.forward_by_name(name, sd, x, torch.float32)
        rng = float(ref.max() - ref.min())
        span = max(1.0, rng)
        m = resselt_b200.load_from_state_dict(dict(sd)).eval().to(dev)
        with torch.inference_mode():
            y32 = m(x.to(dev)).float().cpu()
            m16 = m.bfloat16()
            xb = x.to(dev, torch.bfloat16)
            y16 = m16(xb).float().cpu()
            plan = m16.plan_for(torch.device(dev), torch.bfloat16)
            plan.force_direct = True
            y16d = m16(xb).float().cpu()
            plan.force_direct = False
        e32 = float((y32 - ref).abs().max()) / span
        psnr = lambda y: float(10 * torch.log10(torch.tensor(rng * rng) / ((y - ref) ** 2).mean().clamp_min(1e-30)))
        rec = dict(stage='model', name=name, err_fp32=e32, psnr_bf16_tc=psnr(y16), psnr_bf16_direct=psnr(y16d),
                   tc_vs_direct=float((y16 - y16d).abs().max()) / span)
        rec['ok'] = e32 <= 1e-4 and rec['psnr_bf16_tc'] >= 50.0
        ok_all &= rec['ok']
        print(json.dumps(rec), flush=True)
    return ok_all


def stage_time():
    import torch
    from resselt_b200.archs import SPAN

    dev = torch.device('cuda:0')
    m = SPAN(feature_channels=48, upscale=2, seed=3).eval().to(dev).bfloat16()
    for (n, h, w) in [(1, 256, 256), (1, 1080, 1920)]:
        x = torch.rand(n, 3, h, w, device=dev).bfloat16()
        with torch.inference_mode():
            for _ in range(3):
                y = m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                y = m(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        mp = n * h * w * 4 / 1e6
        plan = m.plan_for(dev, torch.bfloat16)
        print(json.dumps(dict(stage='time', shape=[n, h, w], ms=ms, out_mp_s=mp / ms * 1e3,
                              tflops=plan.flops(n, h, w) / ms / 1e9)), flush=True)
    return True


STAGES = {'layers': stage_layers, 'models': stage_models, 'time': stage_time}

if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in STAGES:
        sys.exit(0 if STAGES[what]() else 1)
    rc = 0
    for name in STAGES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=420, cwd=ROOT)
            code = p.returncode
        except subprocess.TimeoutExpired:
            code = 'timeout'
        print(json.dumps(dict(stage=name, exit=code, secs=round(time.time() - t0, 1))), flush=True)
        rc |= 0 if code == 0 else 1
    sys.exit(rc)
