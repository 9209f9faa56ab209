#!/usr/bin/env python
"""Fold one `ncu --set full` report into profiles/r02_ncu_summary.json (tracked), the file bench.py
reads `roofline.traffic` and the tensor-pipe activity from.

    python profiles/ncu_to_json.py gpurun_out/x.ncu-rep KEY "what was captured"

KEY is the name bench.py looks up (rank_refine_kernel, rank_gemm_kernel, train_hole_step ...).  For a
report with several kernels (a training step) the DRAM bytes are summed over all launches and the
per-kernel rows are kept.  The current commit is recorded next to the numbers.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'r02_ncu_summary.json')
WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_active_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__cycles_elapsed.avg.per_second': 'sm_clock',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed': 'lsu_wavefronts_pct',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'tc_smem_wavefronts_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'lsu_smem_wavefronts_pct',
    'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed': 'smem_bank_reads_pct',
    'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed': 'smem_bank_writes_pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_throughput_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
    'launch__registers_per_thread': 'registers',
    'smsp__inst_executed.sum': 'warp_instructions',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__grid_size': 'grid',
}
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0, 'msecond': 1e-3,
         'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0, 'Ghz': 1e9, 'Mhz': 1e6, 'cycle/nsecond': 1e9}


def rows_of(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, units = r[0], r[1]
    res = []
    for row in r[2:]:
        d = {'kernel': row[h.index('Kernel Name')][:120]}
        for i, n in enumerate(h):
            if n in WANT and row[i] not in ('', 'n/a'):
                v = float(row[i].replace(',', ''))
                d[WANT[n]] = v * SCALE.get(units[i], 1.0)
        res.append(d)
    return res


def main():
    rep, key = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    rows = rows_of(rep)
    commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(['git', '-C', ROOT, 'status', '--porcelain', '--', 'scikit-kge_b200/csrc'],
                                capture_output=True, text=True).stdout.strip())
    total = sum(r.get('dram_read', 0) + r.get('dram_write', 0) for r in rows)
    entry = {'report': os.path.basename(rep), 'commit': commit + ('+uncommitted csrc changes' if dirty else ''),
             'what': note, 'launches': len(rows)}
    if len(rows) == 1:
        r = rows[0]
        entry.update(r)
        entry['dram_bytes_per_launch'] = total
    else:
        entry['dram_bytes_per_step'] = total
        entry['duration_sum'] = sum(r.get('duration', 0) for r in rows)
        entry['kernels'] = rows
    allv = json.load(open(OUT)) if os.path.exists(OUT) else {}
    allv[key] = entry
    with open(OUT, 'w') as f:
        json.dump(allv, f, indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1)[:1500])


if __name__ == '__main__':
    main()
