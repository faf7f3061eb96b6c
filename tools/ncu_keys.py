#!/usr/bin/env python
"""Print the handful of ncu counters the design decisions in profiles/README.md rest on.
    ncu -i report.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_keys.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'lts__t_sectors_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for r in rows[2:]:
    print('-----')
    for w in WANT:
        if w in hdr:
            print(w, '=', r[hdr.index(w)])
