#!/bin/bash
for S in "" "--sorted"; do
  echo "== cutoff sweep $S"
  timeout 300 python tools/kernel_sweep.py $S --case sweep --nx 201 --nrays 2000000 --only traj-tiled256/static/divc,endpoints/refill/divc,endpoints/static/divc,extras/refill/divc,hist64/refill/divc,cross/refill/divc 2>&1 | grep -v "^#"
  echo "== trapped $S"
  timeout 300 python tools/kernel_sweep.py $S --case trapped --nx 257 --nrays 2000000 --only endpoints/refill/divc,extras/refill/divc 2>&1 | grep -v "^#"
done
