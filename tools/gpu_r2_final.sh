#!/bin/bash
# round 2, final verification of this session: smoke(), full GPU suite, long fuzz, bench N=1 both arms, final sweeps
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/final_gpu_tests.txt
FUZZ_ITERS=600 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_rk_gpu.py -m gpu -x -q -k fuzz 2>&1 | tail -2 | tee $O/final_fuzz.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench rc=$?"; tail -3 $O/final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref_n1.json 2> $O/final_bench_ref_n1.err; echo "ref rc=$?"
timeout 400 python tools/kernel_sweep.py --cells both --only endpoints/,extras/,cross/,hist64,hist1024/static/divc/hoist,lookup/,traj-tiled256/static/divc,traj-exact/static/divc,traj-pass2/static/divc,traj-ray/static/divc > $O/final_sweep_config3.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,endpoints/refill/divc,endpoints/static/divc,extras/refill/divc,hist64/refill/divc > $O/final_sweep_trapped.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only traj-exact/static/divc,traj-pass2/static/divc,traj-tiled256/static/divc,endpoints/refill/divc,endpoints/static/divc > $O/final_sweep_cutoff.txt 2>&1
timeout 600 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --cells both --only traj-tiled256/static/divc/cells,endpoints/static/divc,endpoints/refill/divc,extras/static/divc > $O/final_sweep_nx2001.txt 2>&1
timeout 300 python tools/rk_bench.py > $O/final_rk_bench.jsonl 2>&1; tail -4 $O/final_rk_bench.jsonl | cut -c1-400
