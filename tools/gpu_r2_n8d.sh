#!/bin/bash
# round 2, last 8-GPU call: full bench N=8 with the final code (vector unshuffle, 64-register histogram pass)
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8_d.json 2> $O/bench_n8_d.err; echo "bench rc=$?"; tail -2 $O/bench_n8_d.err
