#!/bin/bash
# One GPU session: parity tests, kernel sweep, bench, then the ncu captures (each only after the
# same command exited 0 without ncu).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/kernel_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep.log
timeout 600 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 > gpurun_out/sweep_cutoff.log 2>&1; echo "sweep2 rc=$?"; cat gpurun_out/sweep_cutoff.log
timeout 900 python bench.py --steps 10 --warmup 3 --extra > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
for V in traj-tiled/refill/divc traj-ray/refill/divc endpoints/refill/divc; do
  T=$(echo $V | tr '/' '_')
  SCMD="python tools/kernel_sweep.py --nrays 250000 --reps 1 --only $V"
  timeout 300 $SCMD > gpurun_out/plain_$T.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/prof_$T $SCMD > gpurun_out/ncu_$T.log 2>&1
  echo "ncu $V rc=$?"
done
ls -la gpurun_out
