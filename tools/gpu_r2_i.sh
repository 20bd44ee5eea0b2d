#!/bin/bash
# round 2, call I: lean fast loop (endpoint mode) -- parity suite, then sweeps
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/kernel_sweep.py --hoist both --only endpoints/ > $O/lean_sweep_config3.txt 2>&1; cat $O/lean_sweep_config3.txt
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc,endpoints/static/divc > $O/lean_sweep_trapped.txt 2>&1; cat $O/lean_sweep_trapped.txt
timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only endpoints/refill/divc,endpoints/static/divc > $O/lean_sweep_cutoff.txt 2>&1; cat $O/lean_sweep_cutoff.txt
timeout 300 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --only endpoints/static/divc/hoist,endpoints/static/divc/cells > $O/lean_sweep_nx2001.txt 2>&1; cat $O/lean_sweep_nx2001.txt
