#!/bin/bash
# round 2, call J: y-major cell table -- parity, then x- vs y-major sweeps
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ymajor or hoisted or fan_fixture or guard_special or fuzz or heavy_tail" 2>&1 | tail -5
timeout 600 python tools/kernel_sweep.py --cells both --only endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist,cross/static/divc/hoist,lookup/static/divc/hoist,hist64/static/divc/hoist,traj-tiled256/static/divc/cells,endpoints/static/divc/cells > $O/ymajor_sweep_config3.txt 2>&1; cat $O/ymajor_sweep_config3.txt
timeout 300 python tools/kernel_sweep.py --cells both --case sweep --nx 201 --nrays 2000000 --only endpoints/refill/divc,endpoints/static/divc > $O/ymajor_sweep_cutoff.txt 2>&1; cat $O/ymajor_sweep_cutoff.txt
timeout 300 python tools/kernel_sweep.py --cells both --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc > $O/ymajor_sweep_trapped.txt 2>&1; cat $O/ymajor_sweep_trapped.txt
timeout 300 python tools/kernel_sweep.py --cells both --nx 2001 --nrays 1000000 --only endpoints/static/divc/hoist,endpoints/static/divc/cells > $O/ymajor_sweep_nx2001.txt 2>&1; cat $O/ymajor_sweep_nx2001.txt
STB_HOIST_MAX_MB=256 timeout 300 python tools/kernel_sweep.py --cells both --nx 2001 --nrays 1000000 --only endpoints/static/divc/hoist > $O/ymajor_sweep_nx2001_hoist128MB.txt 2>&1; cat $O/ymajor_sweep_nx2001_hoist128MB.txt
