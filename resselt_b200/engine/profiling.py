"""Device-side timing of a plan and of its launch units, free of host launch cost.

A SPAN 1080p forward is a dozen kernels of 60-170 us; launched one by one from Python through ctypes the host is the
bottleneck on a slow box (round 1's `kernel_ms` read 0.101 ms where the kernel takes 0.078 ms).  Everything here captures
the launches into ONE CUDA graph of ``reps`` repetitions and times the replay with CUDA events, so the number is what the
device spends — and the per-unit times add up to (at most) the whole forward's time.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from .plan import Plan


def time_in_graph(fn: Callable[[], None], reps: int, device: torch.device, replays: int = 3) -> float:
    """Milliseconds per call of ``fn`` (which must only enqueue work on the current stream): ``reps`` calls are captured
    into one CUDA graph; the graph is replayed once untimed and ``replays`` times timed; the best replay is returned."""
    fn()  # binds workspaces / tensor maps outside the capture
    torch.cuda.synchronize(device)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize(device)
    best = float('inf')
    for _ in range(replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def time_forward(plan: Plan, x: torch.Tensor, out: torch.Tensor, reps: int = 10) -> float:
    """Milliseconds per whole forward, device-resident, replayed from a graph."""
    return time_in_graph(lambda: plan.forward(x, out=out), reps, plan.device)


def time_units(plan: Plan, x: torch.Tensor, out: torch.Tensor, reps: int = 20) -> List[Dict]:
    """One record per launch unit of the plan (an op, or a fused pair of ops): kernel name, in-graph milliseconds per launch
    group, algorithmic FLOPs and HBM bytes.  The unit re-runs on whatever the preceding units left in the plan's buffers, so
    its memory traffic is the real one (every activation map of the benchmarked shapes exceeds the 126 MB L2 ... or not: the
    caller states which)."""
    plan.forward(x, out=out)  # fill every buffer, bind the shape
    torch.cuda.synchronize(plan.device)
    records = []
    for begin, end, info in plan.launch_units():
        ms = time_in_graph(lambda b=begin, e=end: plan.forward(x, out=out, ops=(b, e)), reps, plan.device, replays=2)
        records.append(dict(ops=(begin, end), kernel=info['kernel'], launches=info['launches'], ms=ms, flops=info['flops'], bytes=info['bytes']))
    return records


def summarize_units(records: List[Dict], forward_ms: Optional[float] = None) -> Dict[str, Dict]:
    """Group unit records by kernel: launches, total ms, share of the summed unit time, achieved TFLOP/s and GB/s."""
    total = sum(r['ms'] for r in records) or 1.0
    out: Dict[str, Dict] = {}
    for r in records:
        k = out.setdefault(r['kernel'], dict(units=0, ms=0.0, flops=0.0, bytes=0.0))
        k['units'] += 1
        k['ms'] += r['ms']
        k['flops'] += r['flops']
        k['bytes'] += r['bytes']
    for k in out.values():
        k['share'] = k['ms'] / total
        k['tflops'] = k['flops'] / (k['ms'] * 1e-3) / 1e12 if k['ms'] > 0 else 0.0
        k['gbs'] = k['bytes'] / (k['ms'] * 1e-3) / 1e9 if k['ms'] > 0 else 0.0
    if forward_ms is not None:
        out['_forward'] = dict(ms=forward_ms, sum_of_units_ms=total)
    return out
