#!/bin/bash
# round 2, last 8-GPU call: full bench N=8 with the final code (histogram butterfly path, two trace streams, 8 chunks)
set -x
O=gpurun_out/r02
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8_c.json 2> $O/bench_n8_c.err; echo "bench rc=$?"; cat $O/bench_n8_c.json | cut -c1-300; tail -3 $O/bench_n8_c.err
timeout 600 $TR tools/config5_multi_gpu.py --nrays 100000000 --check > $O/config5_n8_c.json 2> $O/config5_n8_c.err; echo "config5 rc=$?"; tail -1 $O/config5_n8_c.json
