#!/bin/bash
# round 2, call K: where does the gather wait of the endpoint kernel come from?  ncu metric subsets, x- vs y-major
set -x
O=gpurun_out/r02
mkdir -p $O
M=l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,dram__bytes_read.sum,gpu__time_duration.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ymajor" 2>&1 | tail -3
for C in x y; do
  for V in endpoints/static/divc/hoist endpoints/static/divc/cells; do
    T=$(echo $V | tr '/' '_')_$C
    SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --cells $C --only $V"
    timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:trace_kernel -s 2 -c 1 --csv --log-file $O/gather_$T.csv $SCMD > $O/gather_$T.log 2>&1
    echo "== $T rc=$?"; grep -E "hit_rate|long_scoreboard|issue_active|lookup_miss|time_duration|dram__bytes|wavefronts" $O/gather_$T.csv | cut -d, -f13-
  done
done
