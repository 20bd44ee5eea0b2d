#!/bin/bash
# round 2, call R: tabulated-equilibrium X mode (STB_RK_MODE_X_TABLE) -- RK suite; full parity suite after the prefetch clean-up
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_rk_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_rk_gpu.py 2>&1 | tail -4
timeout 300 python tools/kernel_sweep.py --cells both --only endpoints/static/divc/hoist,endpoints/refill/divc/hoist,cross/static/divc/hoist 2>&1 | tee $O/after_cleanup_sweep.txt
timeout 300 python tools/rk_bench.py 2>&1 | tail -12
