#!/bin/bash
# One GPU session (run under gpurun): parity tests, smoke, bench (both arms), then the two ncu passes
# of B200_PROFILING.md -- each only after the same command line exited 0 without ncu.
# Outputs under gpurun_out/; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --extra > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traj-export"
timeout 600 $BCMD > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-traj-export"
timeout 600 $BCMD > gpurun_out/bench_plain1.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench $BCMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu full rc=$?"
