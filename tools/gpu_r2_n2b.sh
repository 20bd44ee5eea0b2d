#!/bin/bash
# round 2, 2-GPU call: vectorised unshuffle -- unit test, C two-rank test, config-5 check, config-5 bench block
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank or unshuffle" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/config5_multi_gpu.py --nrays 25000000 --check 2>/dev/null | tail -1
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-traj-export --no-app --no-cpu-baseline --no-parity > $O/bench_n2_b.json 2> $O/bench_n2_b.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/bench_n2_b.json'));print({k:d['config5'].get(k) for k in ('value','ms_per_step','kernel_ms','kernel_ms_fastest_rank','exchange_exposed_ms','wait_after_own_trace_ms_max_rank')})"
