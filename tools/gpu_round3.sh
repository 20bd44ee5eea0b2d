#!/bin/bash
# occupancy (launch bounds) x field layout x schedule experiment
set -u
mkdir -p gpurun_out
ONLY="traj-tiled/refill/divc,endpoints/refill/divc,endpoints/static/divc,endpoints-tiled/refill/divc,extras/refill/divc,extras-tiled/refill/divc,extras/static/divc"
for LIB in libsidtrace_b200.so exp_minb12.so exp_minb16.so; do
  for LAY in rowmajor blocked; do
    echo "##### $LIB $LAY"
    SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/$LIB timeout 300 python tools/kernel_sweep.py --field-layout $LAY --only "$ONLY" 2>&1 | tee -a gpurun_out/sweep3.log
  done
done
echo "##### cutoff sweep"
for LIB in libsidtrace_b200.so exp_minb12.so; do
  for LAY in rowmajor blocked; do
    SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/$LIB timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --field-layout $LAY --only "$ONLY" 2>&1 | tee -a gpurun_out/sweep3_cutoff.log
  done
done
