#!/bin/bash
# round 2, call V: final measurements of this session -- full GPU suite, bench N=1 (both arms), ncu launch list + --set full of the bench
# kernel and of the two-rays-per-thread endpoint kernel, final sweeps
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/v_gpu_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/v_bench_n1.json 2> $O/v_bench_n1.err; echo "bench rc=$?"
cut -c1-400 $O/v_bench_n1.json; tail -3 $O/v_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/v_bench_ref_n1.json 2> $O/v_bench_ref_n1.err; echo "ref rc=$?"
timeout 400 python tools/kernel_sweep.py --cells both --only endpoints/,extras/,cross/,hist64,lookup/,traj-tiled256/static/divc,traj-exact/static/divc,traj-pass2/static/divc,traj-ray/static/divc > $O/v_sweep_config3.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,endpoints/refill/divc,endpoints/static/divc,extras/refill/divc > $O/v_sweep_trapped.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,traj-tiled256/static/divc,endpoints/refill/divc,endpoints/static/divc > $O/v_sweep_cutoff.txt 2>&1
timeout 600 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --cells both --only traj-tiled256/static/divc/cells,endpoints/static/divc,endpoints/refill/divc,extras/static/divc > $O/v_sweep_nx2001.txt 2>&1
cat $O/v_sweep_config3.txt $O/v_sweep_trapped.txt $O/v_sweep_cutoff.txt $O/v_sweep_nx2001.txt
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traj-export --no-config5 --no-app --no-parity"
timeout 600 $BCMD > $O/v_bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/v_launches_bench.csv $BCMD > $O/v_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-traj-export --no-config5 --no-app --no-parity"
timeout 600 $BCMD > $O/v_bench_plain1.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o $O/v_prof_bench $BCMD > $O/v_ncu_bench.log 2>&1
echo "ncu full bench rc=$?"
SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --cells y --only endpoints/static/divc/hoist"
timeout 300 $SCMD > $O/v_plain_dual.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:endpoint_dual_kernel -s 2 -c 1 -f -o $O/v_prof_endpoint_dual $SCMD > $O/v_ncu_dual.log 2>&1
echo "ncu dual rc=$?"
SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --only extras/static/divc/hoist"
timeout 300 $SCMD > $O/v_plain_extras.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o $O/v_prof_extras $SCMD > $O/v_ncu_extras.log 2>&1
echo "ncu extras rc=$?"
ls -la $O | grep " v_"
