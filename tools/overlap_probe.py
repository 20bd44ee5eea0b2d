"""Do the HBM-bound trajectory kernel and the dispatch-bound endpoint kernel overlap when launched on two streams?
(the two passes of trace_exact on different halves of a beam)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayrtrace_b200 import sidtrace, synth
case = synth.beam_case(nx=1001, nrays=1_000_000)
nf = case.nfield
field = sidtrace.Field(np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay), nf.nx, nf.ny)
d_in = [torch.from_numpy(a).cuda() for a in case.launch]
n, L, T = case.nrays, case.L, 256
buf = torch.empty(((n + T - 1) // T, L, T, 4), dtype=torch.float64, device="cuda")
kw = dict(L=L, st=case.st, dx=case.dx, dy=case.dy)
o1 = sidtrace.BatchResult(nsteps=torch.empty(n, dtype=torch.int32, device="cuda"), end_state=torch.empty((4, n), dtype=torch.float64, device="cuda"))
o2 = sidtrace.BatchResult(nsteps=torch.empty(n, dtype=torch.int32, device="cuda"), end_state=torch.empty((4, n), dtype=torch.float64, device="cuda"))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def traj(stream=None): sidtrace.trace_batch(field, *d_in, traj=buf, tiled=True, tile=T, out=o1, stream=stream, **kw)
def endp(stream=None): sidtrace.trace_batch(field, *d_in, out=o2, refill=False, cells="y", stream=stream, **kw)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
def both(first):
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(main); sa.wait_event(ev); sb.wait_event(ev)
    if first == "traj":
        traj(sa); endp(sb)
    else:
        endp(sb); traj(sa)
    ea, eb = torch.cuda.Event(), torch.cuda.Event(); ea.record(sa); eb.record(sb); main.wait_event(ea); main.wait_event(eb)
print("trajectory kernel alone      %.3f ms" % timeit(traj))
print("endpoint kernel alone        %.3f ms" % timeit(endp))
print("both, two streams (traj 1st) %.3f ms" % timeit(lambda: both("traj")))
print("both, two streams (endp 1st) %.3f ms" % timeit(lambda: both("endp")))
