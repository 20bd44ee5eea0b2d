#!/bin/bash
for lib in b200 me8 me10; do
  echo "== $lib"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 400 python tools/kernel_sweep.py --cells y --only extras/static/divc/hoist,extras/refill/divc/hoist,hist64/static/divc/hoist,hist64/refill/divc/hoist 2>&1 | grep -v "^#"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 400 python tools/kernel_sweep.py --nx 201 --nrays 12500000 --L 512 --only hist64/static/divc/hoist,hist64/refill/divc/hoist,extras/static/divc/hoist 2>&1 | grep -v "^#"
done
