#!/bin/bash
# round 2, call A: parity suite + first sweeps of the hoisted kernels / exact arenas / store pattern
set -x
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_a.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_a.log 2>&1; echo "pytest rc=$?" >> $O/pytest_a.log
tail -5 $O/pytest_a.log
timeout 300 python tools/kernel_sweep.py --hoist both --only endpoints/,extras/,cross/,hist64,lookup/ > $O/sweep_beam_modes.txt 2>&1
timeout 300 python tools/kernel_sweep.py --hoist both --only traj-tiled256/static/divc,traj-exact/static/divc,traj-pass2/static/divc > $O/sweep_beam_traj.txt 2>&1
timeout 300 python tools/kernel_sweep.py --store-pattern > $O/store_pattern.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --only traj-exact/static/divc/hoist,traj-pass2/static/divc/hoist,endpoints/refill/divc > $O/sweep_trapped.txt 2>&1
timeout 300 python tools/kernel_sweep.py --case sweep --nx 201 --nrays 2000000 --only traj-exact/static/divc/hoist,traj-pass2/static/divc/hoist,traj-tiled256/static/divc/hoist,endpoints/refill/divc,endpoints/static/divc > $O/sweep_cutoff.txt 2>&1
cat $O/sweep_beam_modes.txt $O/sweep_beam_traj.txt $O/store_pattern.txt $O/sweep_trapped.txt $O/sweep_cutoff.txt | tail -80
