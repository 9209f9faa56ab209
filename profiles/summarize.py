#!/usr/bin/env python
"""Summarise an ncu launch list (CSV, gpu__time_duration.sum) and, optionally, one
`ncu --set full` report into a small text file under profiles/.

    python profiles/summarize.py gpurun_out/launches_r01.csv [gpurun_out/prof.ncu-rep] > profiles/NAME.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_tensor_cycles_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum ', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg ', 'smsp__inst_executed.sum ', 'sm__pipe_tensor_subpipe', 'sm__inst_executed_pipe_tc',
        'smsp__average_warp', 'smsp__warp_issue_stalled']


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.DictReader(io.StringIO(''.join(rows))))
    agg = collections.OrderedDict()
    for x in r:
        k = x['Kernel Name']
        v = float(x['Metric Value'].replace(',', ''))
        a = agg.setdefault(k, [0, 0.0, x['Grid Size'], x['Block Size']])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print('# launch list: %s  (%d launches, %.3f ms of kernel time; ncu times are cold-cache and serialised:' % (
        path, len(r), tot / 1e6))
    print('#  compare SHARES with the CUDA-event numbers in BENCH, not absolutes)')
    print('%-100s %6s %12s %7s  %s' % ('kernel', 'n', 'total ms', 'share', 'grid/block of first launch'))
    for k, (n, v, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print('%-100s %6d %12.3f %6.1f%%  %s %s' % (k[:100], n, v / 1e6, 100 * v / tot, g, b))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    if len(r) < 3:
        print('# (could not read %s)' % path)
        return
    h, units = r[0], r[1]
    for row in r[2:]:
        name = row[h.index('Kernel Name')] if 'Kernel Name' in h else '?'
        print('\n# ncu --set full: %s\n# kernel: %s' % (path, name[:160]))
        for i, n in enumerate(h):
            if any(n.startswith(k.strip()) if k.endswith(' ') else (k in n) for k in KEYS):
                print('%-80s %-14s %s' % (n, units[i], row[i]))


if __name__ == '__main__':
    launches(sys.argv[1])
    for p in sys.argv[2:]:
        full(p)
