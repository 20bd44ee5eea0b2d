#!/bin/bash
# round 2, 2-GPU call: C-ABI comm test, config-5 check, bench N=2
set -x
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank" > $O/pytest_n2.log 2>&1; tail -3 $O/pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/config5_multi_gpu.py --nrays 25000000 --check > $O/config5_n2.json 2> $O/config5_n2.err; echo "config5 rc=$?"; tail -2 $O/config5_n2.json; tail -3 $O/config5_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-traj-export --no-app > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; cat $O/bench_n2.json; tail -3 $O/bench_n2.err
