#!/bin/bash
O=gpurun_out/r02
for i in 1 2; do
timeout 300 python tools/kernel_sweep.py --case trapped --nx 257 --nrays 2000000 --reps 9 --only traj-exact/static/divc,traj-pass2/static/divc,endpoints/refill/divc 2>&1 | grep -v "^# case"
timeout 300 python tools/kernel_sweep.py --reps 9 --only traj-exact/static/divc,traj-pass2/static/divc 2>&1 | grep -v "^# case"
done
python - <<'PY'
import torch, numpy as np, time
from rayrtrace_b200 import sidtrace, synth
case = synth.trapped_case(nx=257, nrays=2000000)
nf = case.nfield
field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
d_in = [torch.from_numpy(a).cuda() for a in case.launch]
kw = dict(L=case.L, st=case.st, dx=case.dx, dy=case.dy)
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r1 = sidtrace.trace_batch(field, *d_in, **kw); torch.cuda.synchronize(); t1 = time.perf_counter()
    pl = sidtrace.plan_trajectories(r1.nsteps, tile=256, L=case.L); torch.cuda.synchronize(); t2 = time.perf_counter()
    arena = torch.empty((pl.total_rows, 256, 4), dtype=torch.float64, device="cuda")
    sidtrace.trace_batch(field, *d_in, traj=arena, plan=pl, want_end=False, **kw); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"pass1 {1e3*(t1-t0):.2f} ms  plan {1e3*(t2-t1):.2f} ms  pass2 {1e3*(t3-t2):.2f} ms")
    del arena
PY
