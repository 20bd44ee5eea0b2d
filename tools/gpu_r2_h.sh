#!/bin/bash
set -x
O=gpurun_out/r02
mkdir -p $O
V="endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist,cross/static/divc/hoist,hist64/static/divc/hoist,lookup/static/divc/hoist"
for lib in b200 pf1 pf2 pf3; do
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --only $V > $O/prefetch_$lib.txt 2>&1
  echo "== $lib"; grep -h "/" $O/prefetch_$lib.txt | grep -v "^#"
done
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_pf2.so timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc/hoist > $O/prefetch_trapped_pf2.txt 2>&1; tail -1 $O/prefetch_trapped_pf2.txt
SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_pf2.so timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only endpoints/refill/divc/hoist,endpoints/static/divc/hoist > $O/prefetch_cutoff_pf2.txt 2>&1; tail -2 $O/prefetch_cutoff_pf2.txt
