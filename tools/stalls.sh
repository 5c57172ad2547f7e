#!/bin/bash
# usage: tools/stalls.sh <lib.so or ''> <kernel regex>   -- prints per-issue stall reasons of one launch (ncu, GPU box)
LIB=$1; K=$2
M=""
for s in wait long_scoreboard short_scoreboard math_pipe_throttle not_selected no_instruction branch_resolving barrier membar sleeping dispatch_stall mio_throttle lg_throttle imc_miss drain tex_throttle; do
  M="$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio"
done
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread$M"
STENOS_B200_LIB=$LIB ncu --metrics $M --clock-control none -k regex:$K -c 1 -s 3 python tools/time_parts.py 2>&1 | grep -E "smsp__|sm__|gpu__|launch__" | sed -e 's/smsp__average_warps_issue_stalled_//' -e 's/_per_issue_active.ratio//'
