#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus 4 --steps 10 --warmup 3 --no-traj-export --no-app > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench rc=$?"; tail -2 $O/bench_n4.err
