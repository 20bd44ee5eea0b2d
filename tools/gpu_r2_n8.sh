#!/bin/bash
# round 2, 8-GPU call: config-5 check at 10^8 rays, bench N=8
set -x
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/config5_multi_gpu.py --nrays 100000000 --check > $O/config5_n8.json 2> $O/config5_n8.err; echo "config5 rc=$?"; tail -2 $O/config5_n8.json; tail -3 $O/config5_n8.err
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"; cat $O/bench_n8.json; tail -3 $O/bench_n8.err
