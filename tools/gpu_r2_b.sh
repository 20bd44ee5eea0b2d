#!/bin/bash
# round 2, call B: L2 policy variants on the trajectory kernel; ncu --set full of endpoints / cross / lookup / bench kernel
set -x
O=gpurun_out/r02
mkdir -p $O
V="traj-tiled256/static/divc,traj-pass2/static/divc,endpoints/static/divc/hoist"
for lib in b200 l2p1 l2p2 l2p3; do
  SIDTRACE_B200_LIB=$PWD/rayrtrace_b200/_lib/libsidtrace_$lib.so timeout 300 python tools/kernel_sweep.py --hoist both --only $V > $O/l2_$lib.txt 2>&1
done
for mb in 20 40; do
  STB_L2_PERSIST_MB=$mb timeout 300 python tools/kernel_sweep.py --hoist both --only $V > $O/l2_persist$mb.txt 2>&1
done
python - <<'PY' > $O/l2_attrs.txt 2>&1
import torch
p = torch.cuda.get_device_properties(0)
print(p)
import ctypes
rt = ctypes.CDLL("libcudart.so.12") if False else None
print("L2", p.L2_cache_size)
PY
grep -h "traj\|endpoints" $O/l2_*.txt
for V in endpoints/static/divc/hoist cross/static/divc/hoist lookup/static/divc/hoist traj-tiled256/static/divc/cells; do
  T=$(echo $V | tr '/' '_')
  SCMD="python tools/kernel_sweep.py --nrays 1000000 --reps 1 --only $V"
  timeout 300 $SCMD > $O/plain_$T.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o $O/prof_$T $SCMD > $O/ncu_$T.log 2>&1
  echo "ncu $V rc=$?"
done
ls -la $O
