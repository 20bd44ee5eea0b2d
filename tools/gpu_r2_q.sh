#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_b200.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ymajor or hoisted or fan_fixture or fuzz or heavy_tail" 2>&1 | tail -2
V="endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist,cross/static/divc/hoist,lookup/static/divc/hoist,hist64/static/divc/hoist"
for lib in pfn0 b200; do
  echo "== $lib config3"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --only $V > $O/touch2_$lib.txt 2>&1
  grep -h "/" $O/touch2_$lib.txt | grep -v "^#"
  echo "== $lib cutoff / trapped"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --case sweep --nx 201 --nrays 2000000 --only endpoints/refill/divc,endpoints/static/divc > $O/touch2_cutoff_$lib.txt 2>&1; grep -h "/" $O/touch2_cutoff_$lib.txt | grep -v "^#"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc > $O/touch2_trapped_$lib.txt 2>&1; grep -h "/" $O/touch2_trapped_$lib.txt | grep -v "^#"
done
