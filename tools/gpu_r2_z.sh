#!/bin/bash
O=gpurun_out/r02
for lib in nobf b200; do
  echo "== $lib"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 600 python tools/kernel_sweep.py --nx 201 --nrays 12500000 --L 512 --only hist64/static/divc/hoist,hist64/refill/divc/hoist 2>&1 | grep -v "^#"
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 600 python tools/kernel_sweep.py --only hist64/static/divc/hoist,hist1024/static/divc/hoist 2>&1 | grep -v "^#"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hist or deposition or config2 or config4 or fuzz" 2>&1 | tail -2
