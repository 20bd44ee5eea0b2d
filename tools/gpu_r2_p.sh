#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --cells y --only endpoints/static/divc/hoist"
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_b200.so timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o $O/prof3_touch_y $SCMD > $O/ncu3_touch.log 2>&1
echo rc=$?
