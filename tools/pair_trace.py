"""Bring-up: per-row clock stamps of the fused conv-pair kernel (CTA 1).  Needs the bring-up library:
    RSB_BRINGUP=1 python -c "import __graft_entry__ as g; g.build()"   (here)
    RSB_BRINGUP=1 python tools/pair_trace.py [act gate]                 (GPU box)
Prints, per row, cycles relative to the first stamp: MMA A (wait full | issued), MMA B (tempty | ofull | issued),
A epilogue (tfull | slot handed back | oempty | math+stores done | ofull), B epilogue (tfull | slot handed back | done)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from resselt_b200.engine import native as N
from tools.pair_check import build

if __name__ == '__main__':
    act = int(sys.argv[1]) if len(sys.argv) > 1 else N.ACT_SILU
    gate = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
    plan, _ = build(act, gate)
    x = torch.randn(1, 48, 1080, 1920).to('cuda:0', torch.bfloat16)
    plan.forward(x)
    for _ in range(3):
        plan.forward(x, ops=(1, 3))
    torch.cuda.synchronize()
    rows = 96
    buf = (C.c_longlong * (4 * rows * 8))()
    n = N.lib().rsb_debug_pair_trace(buf, len(buf))
    assert n == len(buf), n
    t = [[[buf[(role * rows + r) * 8 + k] for k in range(8)] for r in range(rows)] for role in range(4)]
    t0 = min(v for role in t for r in role for v in r if v > 0)
    rel = lambda v: (v - t0) if v > 0 else -1
    print('row | mmaA: wait full issued | mmaB: start tempty-ok ofull-ok issued | epiA: start tfull slot-back pre-oempty oempty math ofull | epiB: start tfull res slot-back done')
    for r in range(rows):
        a, b, ea, eb = t[0][r], t[1][r], t[2][r], t[3][r]
        print(f'{r:3d} | ' + ' '.join(f'{rel(v):7d}' for v in a[:3]) + ' | ' + ' '.join(f'{rel(v):7d}' for v in b[:4]) + ' | ' +
              ' '.join(f'{rel(v):7d}' for v in ea[:7]) + ' | ' + ' '.join(f'{rel(v):7d}' for v in eb[:5]))
