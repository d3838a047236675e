"""Turn ncu outputs brought back in gpurun_out/ into the small text/JSON summaries kept under profiles/.

    python tools/summarize_profile.py launches <launches.csv> <out.md> [title]
    python tools/summarize_profile.py full <report.ncu-rep> <out.md> [title]
    python tools/summarize_profile.py rows <report.ncu-rep> <out.md> [title]
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Grid Size', 'Block Size'))
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        ns = float(r[vi].replace(',', ''))
        name = r[ki].replace('rsb::<unnamed>::', '').split('(')[0]
        key = (name, r[gi], r[bi])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised; '
                'compare shares, not absolutes).\n\n')
        f.write('| kernel | grid | block | launches | total us | avg us | share |\n|---|---|---|---:|---:|---:|---:|\n')
        for (name, grid, block), (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{name}` | {grid} | {block} | {cnt} | {ns / 1e3:.1f} | {ns / 1e3 / cnt:.1f} | {100 * ns / total:.1f}% |\n')
        f.write(f'\nTotal of listed launches: {total / 1e6:.3f} ms over {len(rows) - 1} launches.\n')


WANT = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic',
]


def full(path, out, title):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` on `{os.path.basename(path)}` '
                '(report itself is scratch; these are the numbers quoted in DESIGN.md).\n\n')
        names = [r[ki].replace('rsb::<unnamed>::', '').split('(')[0] for r in rows[2:]]
        f.write('| metric | unit | ' + ' | '.join(f'`{n}`' for n in names) + ' |\n|---|---|' + '---:|' * len(names) + '\n')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f'| {w} | {units[i]} | ' + ' | '.join(r[i] for r in rows[2:]) + ' |\n')
    traffic = {}
    if 'dram__bytes_read.sum' in hdr:
        ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
        for n, r in zip(names, rows[2:]):
            traffic[n] = float(r[ri]) * mult.get(units[ri], 1.0) + float(r[wi]) * mult.get(units[wi], 1.0)
    return traffic


def rows_mode(path, out, title):
    """One row per captured launch (kernels as rows): what a bandwidth-bound kernel needs to be judged."""
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
                                     'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
                                     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
                                     'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')}
    mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
    tmult = {'us': 1.0, 'ms': 1e3, 'ns': 1e-3, 's': 1e6}
    with open(out, 'w') as f:
        f.write(f'# {title}\n\nSource: `ncu --set full --clock-control none --import-source on` on `{os.path.basename(path)}` (one row per captured '
                'launch; cold-cache, serialised). GB/s = (DRAM read + write) / duration; HBM peak in MEASURED_PEAKS.json: 6546 GB/s.\n\n')
        f.write('| kernel | grid x block | us | DRAM read MB | DRAM write MB | GB/s | DRAM % of peak | tensor pipe % | MUFU pipe % | warps active % | issue active % | regs |\n')
        f.write('|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n')
        for r in rows[2:]:
            name = r[col['Kernel Name']].replace('rsb::<unnamed>::', '').split('(')[0]
            us = float(r[col['gpu__time_duration.sum']]) * tmult.get(units[col['gpu__time_duration.sum']], 1.0)
            rd = float(r[col['dram__bytes_read.sum']]) * mult.get(units[col['dram__bytes_read.sum']], 1.0)
            wr = float(r[col['dram__bytes_write.sum']]) * mult.get(units[col['dram__bytes_write.sum']], 1.0)
            g = lambda k: float(r[col[k]])
            f.write(f"| `{name}` | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / us / 1e3:.0f} | "
                    f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                    f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {r[col['launch__registers_per_thread']]} |\n")


if __name__ == '__main__':
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(src)
    if mode == 'launches':
        launches(src, dst, title)
    elif mode == 'rows':
        rows_mode(src, dst, title)
    else:
        t = full(src, dst, title)
        print(json.dumps(t, indent=1))
