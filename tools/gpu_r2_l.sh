#!/bin/bash
# round 2, call L: neighbour prefetch (STB_PF_CELLS = 0..4) in x- and y-major tables
set -x
O=gpurun_out/r02
mkdir -p $O
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_b200.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ymajor or hoisted or fan_fixture or guard_special or fuzz or heavy_tail or canaries" 2>&1 | tail -3
V="endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist,cross/static/divc/hoist,lookup/static/divc/hoist,hist64/static/divc/hoist,endpoints/static/divc/cells"
for lib in pfn0 pfn1 b200 pfn3 pfn4; do
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --only $V > $O/pfn_$lib.txt 2>&1
  echo "== $lib"; grep -h "/" $O/pfn_$lib.txt | grep -v "^#"
done
for lib in pfn0 b200; do
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --case sweep --nx 201 --nrays 2000000 --only endpoints/refill/divc,endpoints/static/divc > $O/pfn_cutoff_$lib.txt 2>&1; grep -h "/" $O/pfn_cutoff_$lib.txt | grep -v "^#"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc > $O/pfn_trapped_$lib.txt 2>&1; grep -h "/" $O/pfn_trapped_$lib.txt | grep -v "^#"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --cells both --nx 2001 --nrays 1000000 --only endpoints/static/divc/hoist,endpoints/static/divc/cells > $O/pfn_nx2001_$lib.txt 2>&1; grep -h "/" $O/pfn_nx2001_$lib.txt | grep -v "^#"
done
