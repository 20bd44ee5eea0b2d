import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rayrtrace_b200 import sidtrace, synth
c = synth.sweep_case(nx=201, nrays=25001)
nf = c.nfield
for rep in range(3):
    t0 = time.perf_counter(); f = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny); t1 = time.perf_counter()
    r = sidtrace.trace_batch(f, *c.launch, L=c.L, st=c.st, dx=c.dx, dy=c.dy); t2 = time.perf_counter()
    f.close(); t3 = time.perf_counter()
    print(f"create {t1-t0:.4f}  trace {t2-t1:.4f}  close {t3-t2:.4f}")
