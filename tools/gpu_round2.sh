#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --extra > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $BCMD > gpurun_out/bench_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_bench_tiled $BCMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"; tail -3 gpurun_out/ncu_bench.log
ls -la gpurun_out | head -40
