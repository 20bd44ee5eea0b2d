#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for V in endpoints/static/divc extras/static/divc hist64/static/divc; do
  T=$(echo $V | tr '/' '_')
  SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --only $V"
  timeout 300 $SCMD > gpurun_out/plain_$T.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/prof2_$T $SCMD > gpurun_out/ncu2_$T.log 2>&1
  echo "ncu $V rc=$?"
done
