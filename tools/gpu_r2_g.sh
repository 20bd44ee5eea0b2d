#!/bin/bash
# round 2, call G: final-kernel measurements -- bench N=1 (both arms), ncu launch list + --set full of the bench kernel,
# endpoints / cross / lookup captures, 2001^2 sweeps with L2 hit rates, heavy-tailed sweeps
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1_final.json 2> $O/bench_n1_final.err; echo "bench rc=$?"
cut -c1-600 $O/bench_n1_final.json; tail -3 $O/bench_n1_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1_final.json 2> $O/bench_ref_n1_final.err; echo "ref rc=$?"
# sweeps (final kernels)
timeout 300 python tools/kernel_sweep.py --hoist both --only endpoints/,extras/,cross/,hist64,lookup/,traj-tiled256/static/divc,traj-exact/static/divc,traj-pass2/static/divc,traj-ray/static/divc > $O/final_sweep_config3.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,endpoints/refill/divc,endpoints/static/divc,extras/refill/divc > $O/final_sweep_trapped.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,traj-tiled256/static/divc,endpoints/refill/divc,endpoints/static/divc > $O/final_sweep_cutoff.txt 2>&1
# 2001^2: config 3's second grid (64 MB of 16-byte cells; the hoisted table would be 128 MB)
timeout 600 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --hoist both --only traj-tiled256/static/divc/cells,endpoints/static/divc,endpoints/refill/divc,extras/static/divc > $O/final_sweep_nx2001.txt 2>&1
STB_HOIST_MAX_MB=256 timeout 600 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --only endpoints/static/divc/hoist,extras/static/divc/hoist > $O/final_sweep_nx2001_hoist128MB.txt 2>&1
cat $O/final_sweep_config3.txt $O/final_sweep_trapped.txt $O/final_sweep_cutoff.txt $O/final_sweep_nx2001.txt $O/final_sweep_nx2001_hoist128MB.txt
# ncu: launch list of the bench command, then --set full of its kernel
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-traj-export --no-config5 --no-app --no-parity"
timeout 600 $BCMD > $O/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv $BCMD > $O/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
BCMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-traj-export --no-config5 --no-app --no-parity"
timeout 600 $BCMD > $O/bench_plain1.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o $O/prof_bench $BCMD > $O/ncu_bench.log 2>&1
echo "ncu full bench rc=$?"
for V in endpoints/static/divc/hoist cross/static/divc/hoist lookup/static/divc/hoist; do
  T=$(echo $V | tr '/' '_')
  SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --only $V"
  timeout 300 $SCMD > $O/plain2_$T.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o $O/prof2_$T $SCMD > $O/ncu2_$T.log 2>&1
  echo "ncu $V rc=$?"
done
SCMD="python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --reps 1 --only endpoints/static/divc/cells --hoist off"
timeout 300 $SCMD > $O/plain2_nx2001.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:trace_kernel -s 2 -c 1 -f -o $O/prof2_endpoints_nx2001_cells $SCMD > $O/ncu2_nx2001.log 2>&1
echo "ncu nx2001 rc=$?"
ls -la $O | tail -30
