#!/bin/bash
# round 2, call D: restructured kernel -- parity suite, then cross/lookup/extras/endpoints sweeps with STB_MINB_CROSS variants
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_d.log 2>&1; echo "pytest rc=$?" >> $O/pytest_d.log
tail -15 $O/pytest_d.log
V="endpoints/static/divc/hoist,endpoints/refill/divc/hoist,extras/static/divc/hoist,cross/static/divc/hoist,cross/refill/divc/hoist,lookup/static/divc/hoist,hist64/static/divc/hoist,traj-tiled256/static/divc/cells,traj-pass2/static/divc/cells"
for lib in b200 mc4 mc5 mc6; do
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --hoist both --only $V > $O/restruct_$lib.txt 2>&1
  echo "== $lib"; grep -h "/" $O/restruct_$lib.txt | grep -v "^#"
done
