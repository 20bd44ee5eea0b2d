#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_rays" 2>&1 | tail -3
for S in off timed reuse; do
  echo "== cutoff sweep, sort-rays $S"
  timeout 300 python tools/kernel_sweep.py --sort-rays $S --case sweep --nx 201 --nrays 2000000 --only traj-tiled256/static/divc,endpoints/refill/divc,endpoints/static/divc,extras/refill/divc,cross/refill/divc,hist64/refill/divc 2>&1 | grep -v "^#"
done 2>&1 | tee $O/sorted_rays_raw.txt
