#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block per profiled launch with the metrics that matter
for a streaming kernel (duration, DRAM bytes, DRAM %, FP64 pipe %, issue %, occupancy, stall reasons)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg']
for r in data:
    print('----')
    for w in want:
        if w in idx:
            print(f'{w} = {r[idx[w]][:110]} {units[idx[w]]}')
    stalls = [(float(r[i].replace(',', '')), h) for h, i in idx.items()
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f'  stall {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")}: {v:.2f}')
