#!/bin/bash
# last verification: smoke, the whole GPU suite, bench N=1 (short)
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/final2_gpu_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/final2_bench_n1.json 2> $O/final2_bench_n1.err; echo "bench rc=$?"; tail -2 $O/final2_bench_n1.err
python -c "
import json
d=json.load(open('$O/final2_bench_n1.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'endpoint',d['endpoint_mode']['value'],'exact',d['exact_arena']['value'],'config5',d['config5']['value'],d['config5']['ms_per_step'],'app',d['app']['scan_7x1e6']['wall_s'],'parity',d['parity_check']['result'])
"
