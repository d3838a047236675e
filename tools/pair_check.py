"""Bring-up check of the fused conv-pair kernel (csrc/conv_pair.cu) on a B200: fused pair vs the two separate
row-streaming launches, bit for bit, on several shapes; then timing at 1080p.
Usage (GPU box): python tools/pair_check.py [--time]"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from resselt_b200.engine import native as N
from resselt_b200.engine.plan import INPUT, OUTPUT, PlanBuilder

DEV = 'cuda:0'


def build(act, gate, seed=0):
    g = torch.Generator().manual_seed(seed)
    pb = PlanBuilder(torch.bfloat16, 48, 48, 1)
    a, b, c = pb.buffer(48), pb.buffer(48), pb.buffer(48)
    w1 = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
    w2 = torch.randn(48, 48, 3, 3, generator=g) / (48 * 9) ** 0.5
    pb.conv(INPUT, a, torch.eye(48).view(48, 48, 1, 1))
    pb.conv(a, b, w1, torch.randn(48, generator=g) * 0.1, act=act)
    if gate:
        pb.conv(b, c, w2, torch.randn(48, generator=g) * 0.1, combine=N.COMB_SPAB_GATE, res1=a)
    else:
        pb.conv(b, c, w2, torch.randn(48, generator=g) * 0.1, act=act)
    pb.conv(c, OUTPUT, torch.eye(48).view(48, 48, 1, 1))
    return pb.finalize(torch.device(DEV)), c


def run_case(n, H, W, act, gate):
    plan, c = build(act, gate)
    x = torch.randn(n, 48, H, W, generator=torch.Generator().manual_seed(H * 7 + W)).to(DEV, torch.bfloat16)
    got = {}
    for mode, fd in (('pair', 4), ('rs', 3)):
        plan.force_direct = fd
        plan.forward(x)
        torch.cuda.synchronize()
        got[mode] = plan.read_buffer(c).cpu()
    same = torch.equal(got['pair'], got['rs'])
    diff = (got['pair'] - got['rs']).abs()
    msg = f'n={n} {H}x{W} act={act} gate={gate}: pair==rs {same} (max diff {float(diff.max()):.3e}, launches {plan.launches_per_forward})'
    if not same:
        idx = torch.nonzero(diff > 0)
        msg += f'  mismatches {idx.shape[0]}, first {idx[0].tolist()}, rows {sorted(set(idx[:, 2].tolist()))[:12]}, cols {sorted(set(idx[:, 3].tolist()))[:12]}'
    print(('OK   ' if same else 'FAIL ') + msg, flush=True)
    return same


def time_pair():
    for label, act, gate in (('silu+gate', N.ACT_SILU, True), ('mish+gate', N.ACT_MISH, True), ('silu+silu', N.ACT_SILU, False)):
        plan, _ = build(act, gate)
        x = torch.randn(1, 48, 1080, 1920).to(DEV, torch.bfloat16)
        plan.forward(x)
        for mode, fd in (('pair', 4), ('rs', 3)):
            plan.force_direct = fd
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
            us = e0.elapsed_time(e1) * 1e3 / reps
            fl = 2 * 2 * 48 * 48 * 9 * 1080 * 1920
            print(f'pair 48->48->48 3x3 1080p {label:10s} {mode:4s}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    t0 = time.time()
    ok = True
    for case in [(1, 300, 512, N.ACT_SILU, True), (1, 1200, 128, N.ACT_SILU, True), (2, 400, 248, N.ACT_MISH, True), (1, 1500, 120, N.ACT_NONE, False),
                 (3, 211, 256, N.ACT_SILU, False), (1, 1080, 1920, N.ACT_SILU, True)]:
        ok &= run_case(*case)
    print('ALL OK' if ok else 'SOME FAILED', f'({time.time() - t0:.1f} s)')
    if '--time' in sys.argv:
        time_pair()
    sys.exit(0 if ok else 1)
