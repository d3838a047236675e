"""Summarise a RSB_PAIR_TRACE dump (CTA 1 of the fused pair kernel): per-step phases of the MMA thread and per-row
phases of one A-epilogue / B-epilogue lane, in SM cycles.   python tools/pair_trace_report.py <trace.txt>"""
import sys

import numpy as np

rows = [l.split() for l in open(sys.argv[1])]
d = {0: [], 1: [], 2: []}
for r in rows:
    d[int(r[0])].append([int(v) for v in r[2:]])
m, a, b = (np.array(d[i]) for i in range(3))
t0 = m[0, 0]
print('MMA thread, step k: start | A: wait tempty, wait full, issue | B: wait tempty, wait ofull, issue')
for k in list(range(3, 10)) + list(range(50, 60)):
    r = m[k]
    if r[0] == 0 or r[4] == 0:
        continue
    print(f'{k:3d} {r[0] - t0:7d} | {r[1] - r[0]:5d} {r[2] - r[1]:5d} {r[3] - r[2]:5d} | {r[4] - r[3]:5d} {r[5] - r[4]:5d} {r[6] - r[5]:5d}')
print('A epilogue lane, row q: start | wait tfull, drain, wait oempty, math + st.shared, fence + arrive')
for q in list(range(0, 10, 2)) + list(range(50, 60, 2)):
    r = a[q]
    if r[0] == 0:
        continue
    print(f'{q:3d} {r[0] - t0:7d} | {r[1] - r[0]:5d} {r[2] - r[1]:5d} {r[3] - r[2]:5d} {r[4] - r[3]:5d} {r[5] - r[4]:5d}')
print('B epilogue lane, row q: start | wait tfull, drain, residual + math + store')
for q in list(range(0, 10, 2)) + list(range(50, 60, 2)):
    r = b[q]
    if r[0] == 0:
        continue
    print(f'{q:3d} {r[0] - t0:7d} | {r[1] - r[0]:5d} {r[2] - r[1]:5d} {r[3] - r[2]:5d}')
