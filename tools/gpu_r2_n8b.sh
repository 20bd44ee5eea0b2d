#!/bin/bash
# round 2, second 8-GPU call: config-5 check at 10^8 rays, config-5 block with 4 / 16 chunks, full bench N=8 (8 chunks)
set -x
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/config5_multi_gpu.py --nrays 100000000 --check > $O/config5_n8_b.json 2> $O/config5_n8_b.err; echo "config5 rc=$?"; tail -2 $O/config5_n8_b.json; tail -3 $O/config5_n8_b.err
for C in 4 16; do
  timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --config5-chunks $C --no-cpu-baseline --no-traj-export --no-app --no-parity > $O/bench_n8_chunks$C.json 2> $O/bench_n8_chunks$C.err; echo "bench chunks $C rc=$?"
  python -c "import json;d=json.load(open('$O/bench_n8_chunks$C.json'));print({k:d['config5'][k] for k in ('value','ms_per_step','kernel_ms','exchange_exposed_ms','exchange_frac_of_pass')})"
done
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8_b.json 2> $O/bench_n8_b.err; echo "bench rc=$?"; cat $O/bench_n8_b.json; tail -3 $O/bench_n8_b.err
