#!/usr/bin/env python3
"""Prints the metrics we track from an .ncu-rep (read on the CPU box): python tools/ncu_summary.py REPORT"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct', 'l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg']
if sys.argv[1].endswith('.csv'):      # a raw page exported on the GPU box (tools/ncu_export.sh)
    rows = list(csv.reader(open(sys.argv[1])))
else:
    rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('%-86s %-16s %s' % (k, units[i], r[i][:110]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') or ('warps_issue_stalled' in h and 'per_warp_active' in h and False):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.03:
                print('%-86s %-16s %.3f' % (h, units[i], v))
    print()
