#!/bin/bash
O=gpurun_out/r02
echo "== fixed-cell diagnostic (every gather an L1 hit; results are wrong by design), dual kernel MINB 9"
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_diag.so timeout 300 python tools/kernel_sweep.py --only endpoints/static/divc/hoist,endpoints/static/divc/cells 2>&1 | tee $O/diag_fixed_cell.txt
echo "== normal, dual kernel MINB 9"
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_dual9.so timeout 300 python tools/kernel_sweep.py --only endpoints/static/divc/hoist,endpoints/static/divc/cells 2>&1 | grep -v "^#"
