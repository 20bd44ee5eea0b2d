#!/bin/bash
# round 2, call Y: deposition histogram -- log-step peer reduction, neighbour-bin annuli from shared memory
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hist or deposition or config2 or config4 or fuzz or hoisted or ymajor or canaries" 2>&1 | tail -3
timeout 600 python tools/kernel_sweep.py --nx 201 --nrays 12500000 --L 512 --only hist64/static/divc/hoist,hist64/refill/divc/hoist,hist1024/refill/divc/hoist 2>&1 | tee $O/hist_opt_config5shape.txt
timeout 600 python tools/kernel_sweep.py --only hist64/static/divc/hoist,hist64/refill/divc/hoist,hist1024/refill/divc/hoist,hist1024/static/divc/hoist 2>&1 | tee $O/hist_opt_config3.txt
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only hist64/refill/divc/hoist 2>&1 | tee $O/hist_opt_trapped.txt
