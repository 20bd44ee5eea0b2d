#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
V="extras/static/divc/hoist,cross/static/divc/hoist,lookup/static/divc/hoist,hist64/static/divc/hoist,traj-tiled256/static/divc"
for lib in b200 unr2; do
  echo "== $lib"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 400 python tools/kernel_sweep.py --only $V 2>&1 | grep -v "^#" | tee $O/crossopt_$lib.txt
done
