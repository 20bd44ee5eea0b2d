#!/bin/bash
# final ncu --set full captures of the passes the application runs (crossing + lookup on the persistent schedule, y-major table)
O=gpurun_out/r02; mkdir -p $O
for V in cross/refill/divc/hoist lookup/refill/divc/hoist hist64/static/divc/hoist; do
  T=$(echo $V | tr '/' '_')
  SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --cells y --only $V"
  timeout 300 $SCMD > $O/plain3_$T.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o $O/prof3_$T $SCMD > $O/ncu3_$T.log 2>&1
  echo "ncu $V rc=$?"
done
