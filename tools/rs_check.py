"""Bring-up check of the row-streaming 3x3 kernel (csrc/conv_rs.cu) on a B200:
single layers vs an fp64 reference and vs the tile kernel (bit for bit), then per-layer timing at 1080p.
Usage (GPU box): python tools/rs_check.py [--time]
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from resselt_b200.engine import native as N
from resselt_b200.engine.plan import INPUT, OUTPUT, PlanBuilder

DEV = 'cuda:0'

CASES = [
    # cin, cout, n, H, W, act, gate
    (48, 48, 1, 32, 40, N.ACT_NONE, False),
    (48, 48, 1, 1, 8, N.ACT_NONE, False),
    (48, 48, 1, 2, 136, N.ACT_SILU, False),
    (48, 48, 2, 300, 256, N.ACT_SILU, False),
    (48, 48, 1, 64, 384, N.ACT_NONE, True),
    (48, 48, 3, 7, 128, N.ACT_MISH, False),
    (64, 64, 2, 33, 24, N.ACT_LRELU, False),
    (160, 32, 1, 40, 160, N.ACT_LRELU, False),
    (48, 12, 1, 32, 24, N.ACT_NONE, False),
    (80, 40, 1, 19, 32, N.ACT_SIGMOID, False),
    (48, 48, 1, 1080, 1920, N.ACT_SILU, False),
]


def run_case(cin, cout, n, H, W, act, gate):
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    x = torch.randn(n, cin, H, W, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    pb = PlanBuilder(torch.bfloat16, cin, cout, 1)
    a, b = pb.buffer(cin), pb.buffer(cout)
    pb.conv(INPUT, a, torch.eye(cin).view(cin, cin, 1, 1))
    if gate:
        pb.conv(a, b, wt, bias, combine=N.COMB_SPAB_GATE, res1=a)
    else:
        pb.conv(a, b, wt, bias, act=act, act_param=0.2)
    pb.conv(b, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    plan = pb.finalize(torch.device(DEV))
    xd = x.to(DEV, torch.bfloat16)
    got = {}
    for mode, fd in (('rs', 3), ('tc', 2)):
        plan.force_direct = fd
        plan.forward(xd)
        torch.cuda.synchronize()
        got[mode] = plan.read_buffer(b).cpu()
    same = torch.equal(got['rs'], got['tc'])
    maxdiff = float((got['rs'] - got['tc']).abs().max())
    msg = f'cin={cin} cout={cout} n={n} {H}x{W} act={act} gate={gate}: rs==tc {same} (max diff {maxdiff:.3e})'
    if H * W <= 300 * 256:
        q = lambda t: t.to(torch.bfloat16).double()
        ref = F.conv2d(q(x), q(wt), bias.double(), padding=1)
        if gate:
            ref = (ref + q(x)) * (torch.sigmoid(ref) - 0.5)
        else:
            ref = {N.ACT_NONE: lambda t: t, N.ACT_SILU: F.silu, N.ACT_MISH: F.mish, N.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2),
                   N.ACT_SIGMOID: torch.sigmoid}[act](ref)
        err = float((got['rs'].double() - ref).abs().max()) / float(ref.abs().max())
        msg += f'  rs vs fp64 {err:.3e}'
        ok = same and err < 8e-3
    else:
        ok = same
    print(('OK   ' if ok else 'FAIL ') + msg, flush=True)
    return ok


def time_layers():
    for label, act, gate in (('plain', N.ACT_NONE, False), ('silu', N.ACT_SILU, False), ('mish', N.ACT_MISH, False), ('gate', N.ACT_NONE, True)):
        pb = PlanBuilder(torch.bfloat16, 3, 3, 1)
        a, b = pb.buffer(48), pb.buffer(48)
        g = torch.Generator().manual_seed(1)
        wt = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
        pb.conv(INPUT, a, torch.randn(48, 3, 3, 3, generator=g) * 0.2)
        if gate:
            pb.conv(a, b, wt, torch.zeros(48), combine=N.COMB_SPAB_GATE, res1=a)
        else:
            pb.conv(a, b, wt, torch.zeros(48), act=act)
        pb.conv(b, a, wt, torch.zeros(48), act=act)  # second layer so the pair ping-pongs
        pb.conv(b, OUTPUT, torch.randn(3, 48, 3, 3, generator=g) * 0.1)
        plan = pb.finalize(torch.device(DEV))
        x = torch.rand(1, 3, 1080, 1920, device=DEV, dtype=torch.bfloat16)
        plan.forward(x)
        for mode, fd in (('rs', 3), ('tc', 2)):
            plan.force_direct = fd
            for _ in range(3):
                plan.forward(x, ops=(1, 2))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                plan.forward(x, ops=(1, 2))
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            fl = 2 * 48 * 48 * 9 * 1080 * 1920
            print(f'layer 48->48 3x3 1080p {label:5s} {mode}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    t0 = time.time()
    allok = True
    for c in CASES:
        allok &= run_case(*c)
    print('ALL OK' if allok else 'SOME FAILED', f'({time.time() - t0:.1f} s)')
    if '--time' in sys.argv:
        time_layers()
    sys.exit(0 if allok else 1)
