#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` into the handful of numbers DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'dram__bytes_write.sum.per_second', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__cycles_elapsed.max', 'sm__cycles_elapsed.max.per_second',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_uniform.sum',
    'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
    'l1tex__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_tmem.sum', 'smsp__inst_executed_op_tmem.sum',
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            base = h.split('.TriageCompute.')[-1]
            if h == 'Kernel Name':
                d['kernel'] = v
            elif base in KEEP and v != '':
                d[base] = '%s %s' % (v, u) if u else v
        res.append(d)
    print(json.dumps({'report': rep.split('/')[-1], 'launches': res}, indent=1))


if __name__ == '__main__':
    main()
