#!/bin/bash
for lib in b200 mc6 mc7 mc8; do
  echo "== $lib"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 400 python tools/kernel_sweep.py --cells y --only cross/static/divc/hoist,cross/refill/divc/hoist,lookup/static/divc/hoist,lookup/refill/divc/hoist 2>&1 | grep -v "^#"
done
