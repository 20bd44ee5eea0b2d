#!/usr/bin/env python
"""Per-call latency of the legacy drop-in surface: sidtrace.trace(...) called once per ray exactly as
the reference's ray.py:80-83 does (BASELINE config 1 field, 204-step rays), and the product's
Ray(...) constructor.  Wall clock, one GPU."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayrtrace_b200 import sidtrace, synth  # noqa: E402
from rayrtrace_b200.nfield import Nfield  # noqa: E402
from rayrtrace_b200.ray import Ray  # noqa: E402

n, x, y = synth.shaped_density(201, shift=0.0, elon=1.0)
nf = Nfield(n, x=x, y=y)
ax, ay = np.ascontiguousarray(nf.ax), np.ascontiguousarray(nf.ay)
L = 100000
bufs = [np.zeros(L) for _ in range(4)]
xg0, yg0 = float(nf.x2grid(-3.0)), float(nf.y2grid(-9.5))
k0 = 2 * np.pi * 136.268e9 / 29979245800.0


def one():
    bufs[0][0], bufs[1][0], bufs[2][0], bufs[3][0] = xg0, yg0, 0.0, k0
    return sidtrace.trace(*bufs, nf.st, nf.dx[0], nf.dy[0], nf.nx, nf.ny, ax, ay)


for _ in range(20):
    steps = one()
t0 = time.perf_counter()
N = 2000
for _ in range(N):
    one()
dt = (time.perf_counter() - t0) / N
print(f"sidtrace.trace: {dt * 1e6:.1f} us per call, {steps} steps -> {steps / dt:.3e} ray-steps/s (one ray at a time)")
t0 = time.perf_counter()
M = 200
for _ in range(M):
    r = Ray(-3.0, -9.5, 0, 1, 136.268e9, 1, n_field=nf)
dt = (time.perf_counter() - t0) / M
print(f"Ray(...): {dt * 1e6:.1f} us per ray ({len(r.x)} points; includes np.zeros(4 x {Ray.MaxStep}) like the reference, ray.py:19-22)")
