#!/bin/bash
# round 2, call O: ncu of the touch-load variant (y-major) against no prefetch
O=gpurun_out/r02
mkdir -p $O
M=l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,gpu__time_duration.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for lib in pfn0 b200; do
  for C in y; do
    T=${lib}_$C
    SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --cells $C --only endpoints/static/divc/hoist"
    SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:trace_kernel -s 2 -c 1 --csv --log-file $O/touchncu_$T.csv $SCMD > $O/touchncu_$T.log 2>&1
    echo "== $T rc=$?"; grep -E "hit_rate|stalled|issue_active|lookup_|time_duration|wavefronts|requests|inst_executed" $O/touchncu_$T.csv | awk -F'","' '{print $(NF-2), $NF}'
  done
done
