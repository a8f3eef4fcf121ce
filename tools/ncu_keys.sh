#!/bin/bash
# usage: ncu_keys.sh file.ncu-rep  -- the handful of raw metrics that matter, one per line
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
keys=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy','sm__pipe_fma_cycles_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__average_warp_latency_issue_stalled','smsp__average_warps_issue_stalled','sm__throughput.avg.pct','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','local','smsp__cycles_active.avg','sm__cycles_elapsed.max']
for h,u,v in zip(hdr,units,vals):
    if any(k in h for k in keys): print('%-95s %-14s %s'%(h,u,v))
"
