#!/bin/bash
# round 2, call S: two rays per thread in endpoint mode (endpoint_dual_kernel), occupancy variants
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
V="endpoints/static/divc/hoist,endpoints/static/divc/cells,endpoints/static/ieee/hoist"
echo "== one ray per thread (STB_DUAL=0)"
STB_DUAL=0 timeout 300 python tools/kernel_sweep.py --cells both --only $V 2>&1 | grep -v "^#" | tee $O/dual_off.txt
for lib in dual6 b200 dual9 dual10; do
  echo "== $lib"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --only $V 2>&1 | grep -v "^#" | tee $O/dual_$lib.txt
done
echo "== cutoff / trapped / 2001 (b200 = MINB 8), dual on then off"
for D in 1 0; do
  STB_DUAL=$D timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only endpoints/static/divc 2>&1 | grep -v "^#"
  STB_DUAL=$D timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only endpoints/static/divc 2>&1 | grep -v "^#"
  STB_DUAL=$D timeout 300 python tools/kernel_sweep.py --nx 2001 --nrays 1000000 --cells both --only endpoints/static/divc/hoist 2>&1 | grep -v "^#"
done | tee $O/dual_other_cases.txt
