#!/usr/bin/env python
"""Headline benchmark: output megapixels/s of SPAN 2x on 1080p frames (bf16), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one forward of one 1x3x1080x1920 frame per GPU (weak scaling: every rank upscales its own
frames, no collective on the compute path).  Rank 0 prints ONE JSON line.

  value        device-resident throughput: frames already in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public API with HOST buffers: pinned H2D copy of every frame and
               D2H copy of every upscaled frame inside the timed region (resselt_b200.runner.FramePipeline)
  roofline     the dominant kernel (3x3 48->48 tensor-core conv + SiLU) timed alone with CUDA events
               against the measured dense-bf16 peak (MEASURED_PEAKS.json, burst figure)
  cpu_baseline the CPU oracle (restatement of the reference forward) on this host's cores, bounded sample

--impl reference times the reference's own algorithm on the host CPU (oracle port: the reference is a
pure-Python package over PyTorch ATen and does not travel to the GPU box), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, SCALE, FEATURES = 1080, 1920, 2, 48
OUT_MP = H * W * SCALE * SCALE / 1e6
WORKLOAD = 'SPAN 2x (feature_channels=48) on 1x3x1080x1920 frames -> 3x2160x3840'
METRIC = 'output megapixels/s (SPAN 2x 1080p bf16)'
WEIGHT_SEED = 3


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(source='measured', bf16_burst=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    hbm=d['hbm_gbs'])
    return dict(source='fallback', bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0)


class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
               0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as exc:  # NVML missing: report it, do not fail the bench
            self.reasons.add(f'nvml_unavailable:{type(exc).__name__}')

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


def _oracle_state_dict():
    from resselt_b200.archs import SPAN

    return {k: v.clone() for k, v in SPAN(num_in_ch=3, num_out_ch=3, feature_channels=FEATURES, upscale=SCALE, seed=WEIGHT_SEED).state_dict().items()}


def _cpu_forward_rate(h: int, w: int, steps: int, warmup: int):
    """Oracle (CPU restatement of the reference forward) on all host threads; returns (out MP/s, ms/step, threads)."""
    import torch

    import oracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = _oracle_state_dict()
    x = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        oracle.forward_by_name('SPAN', sd, x, torch.float32)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.forward_by_name('SPAN', sd, x, torch.float32)
    dt = (time.perf_counter() - t0) / steps
    return h * w * SCALE * SCALE / 1e6 / dt, dt * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    # bounded sample per step: a 270x480 crop (1/16 of a 1080p frame) so K+W steps stay within minutes
    sh, sw = 270, 480
    value, ms, threads = _cpu_forward_rate(sh, sw, args.steps, args.warmup)
    sample = f'{args.steps} steps of one {sh}x{sw} crop (1/16 of a 1080p frame), fp32, torch CPU ops, {threads} threads'
    line = dict(
        impl='reference', metric=METRIC, value=value, unit='MP/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
        config=dict(workload=WORKLOAD, sample=sample),
        cpu_baseline=dict(value=value, unit='MP/s', cores=threads, kind='port', sample=sample),
        e2e=dict(value=value, unit='MP/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line), flush=True)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def run_engine(args):
    import torch
    import torch.distributed as dist

    import resselt_b200
    from resselt_b200.runner import FramePipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    try:
        # one process per GPU: run on the CPUs next to that GPU, so that the pinned frame buffers (first touch) and the copy
        # engines' host traffic stay on the GPU's own NUMA node — matters for the end-to-end number when 8 ranks stream at once
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:  # noqa: BLE001 - placement is an optimisation
        pass
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sd = _oracle_state_dict()
    model = resselt_b200.load_from_state_dict(dict(sd)).eval().to(dev).bfloat16()
    plan = model.plan_for(dev, torch.bfloat16)
    n_inputs = 8  # rotate over distinct frames; the per-step working set (1.6 GB of activations) dwarfs the 126 MB L2 anyway
    gen = torch.Generator().manual_seed(100 + rank)
    host_frames = [torch.rand(1, 3, H, W, generator=gen).to(torch.bfloat16).pin_memory() for _ in range(n_inputs)]
    dev_frames = [f.to(dev) for f in host_frames]
    out = torch.empty((1, 3, H * SCALE, W * SCALE), dtype=torch.bfloat16, device=dev)

    sampler = ClockSampler(local) if rank == 0 else None
    # ---------------------------------------------------------------- device-resident throughput
    with torch.inference_mode():
        for i in range(n_inputs):  # untimed priming: every (frame, out) pair is seen once, so its CUDA graph exists before the warm-up
            model.forward_into(dev_frames[i], out)
        for i in range(args.warmup):
            model.forward_into(dev_frames[i % n_inputs], out)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            model.forward_into(dev_frames[i % n_inputs], out)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = world * args.steps * OUT_MP / (ms_total / 1e3)

    # ---------------------------------------------------------------- end to end (host buffers, copies timed)
    pipe = FramePipeline(model, SCALE, dev, depth=3)
    frames = [host_frames[i % n_inputs] for i in range(args.steps)]
    host_out = [torch.empty((1, 3, H * SCALE, W * SCALE), dtype=torch.bfloat16).pin_memory() for _ in range(min(args.steps, 4))]
    outs = [host_out[i % len(host_out)] for i in range(args.steps)]
    pipe.run(frames[: max(3, args.warmup)], outs[: max(3, args.warmup)])
    barrier()
    t0 = time.perf_counter()
    pipe.run(frames, outs)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_value = world * args.steps * OUT_MP / (e2e_ms / 1e3)
    h2d = host_frames[0].numel() * host_frames[0].element_size()
    d2h = host_out[0].numel() * host_out[0].element_size()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- dominant kernel alone (roofline)
    # op 1 of the SPAN plan is block_1.c1_r: the 3x3 48->48 tensor-core conv + SiLU that makes up 12 of the 23
    # launches (its gate / plain siblings share the kernel template); re-run just that op on the live buffers
    peaks = _peaks()
    f = FEATURES
    iters = 50
    with torch.inference_mode():
        for _ in range(3):
            plan.forward(dev_frames[0], out=out, ops=(1, 2))
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(iters):
            plan.forward(dev_frames[0], out=out, ops=(1, 2))
        k1.record()
        torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / iters
    kernel_flops = 2.0 * f * f * 9 * H * W
    achieved = kernel_flops / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get('dram_bytes_per_launch')
    step_tflops = plan.flops(1, H, W) * world * args.steps / (ms_total * 1e-3) / 1e12

    # ---------------------------------------------------------------- CPU baseline (bounded sample, this host)
    cpu_h, cpu_w = 540, 960  # a quarter frame: ~10-30 s of CPU work with one warm-up
    cpu_value, cpu_ms, cpu_threads = _cpu_forward_rate(cpu_h, cpu_w, steps=2, warmup=1)

    line = dict(
        metric=METRIC, value=value, unit='MP/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_total / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
        config=dict(workload=WORKLOAD, frames_per_step_per_gpu=1, sharding='frames round-robin over ranks, no collective',
                    l2='per-step working set (1.6 GB of activations) exceeds the 126 MB L2; inputs rotate over 8 frames',
                    launch='whole forward replayed as a CUDA graph per (frame, out) pair (Plan.forward graph=True); captured in an untimed priming pass',
                    weights='random init (seeded), loaded through resselt_b200.load_from_state_dict'),
        clocks=clocks,
        e2e=dict(value=e2e_value, unit='MP/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / args.steps,
                 api='resselt_b200.runner.FramePipeline over load_from_state_dict(...) module, pinned host frames'),
        gpu_launches=plan.launches_per_forward * args.steps,
        roofline=dict(bound='tensor', achieved=achieved, peak=peaks['bf16_burst'], unit='TFLOP/s', frac=achieved / peaks['bf16_burst'],
                      traffic=traffic, kernel='conv_rs (row-streaming tcgen05) 3x3 48->48 + SiLU, 1080p', kernel_ms=kernel_ms, flops_per_launch=kernel_flops,
                      peak_source=peaks['source'] + ' burst bf16 (kernel timed alone)',
                      step_tflops=step_tflops, step_frac_of_sustained=step_tflops / (peaks['bf16_sustained'] * world)),
        cpu_baseline=dict(value=cpu_value, unit='MP/s', cores=cpu_threads, kind='port',
                          sample=f'2 forwards of one {cpu_h}x{cpu_w} crop (quarter of a 1080p frame), fp32, after 1 warm-up'),
    )
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'engine' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}', '--master-addr', '127.0.0.1',
               '--master-port', str(_free_port()), os.path.abspath(__file__), '--gpus', str(args.gpus), '--steps', str(args.steps),
               '--warmup', str(args.warmup)]
        raise SystemExit(subprocess.call(cmd))
    run_engine(args)


if __name__ == '__main__':
    main()
