#!/bin/bash
# round 2, call M: occupancy curve of the endpoint kernel (how long is one step of one warp when the SM is empty?)
O=gpurun_out/r02
mkdir -p $O
export SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_pfn0.so
for n in 4736 18944 37888 75776 151552 227328 303104 454656 1000000; do
  timeout 300 python tools/kernel_sweep.py --nrays $n --reps 5 --only endpoints/static/divc/hoist,endpoints/static/divc/cells 2>&1 | grep -v "^# case" | sed "s/^/n=$n /"
done | tee $O/occupancy_curve_endpoints.txt
