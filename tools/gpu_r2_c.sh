#!/bin/bash
# round 2, call C: parity suite, smoke, bench N=1 (both arms)
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_c.log 2>&1; echo "pytest rc=$?" >> $O/pytest_c.log
tail -5 $O/pytest_c.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
cat $O/bench_n1.json; tail -5 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"
cat $O/bench_ref_n1.json
