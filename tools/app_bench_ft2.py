#!/usr/bin/env python
"""Application-level timing of the reference's top-level driver shape (ft2.phases, ft2.py:25-71):
density scan x 7 TX/RX antenna pairs x nray rays, on one GPU.  Wall-clock per call, device path
(two O(rays)-output launches per beam) vs host path (materialised Ray objects + addpoint + interp1d)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayrtrace_b200 import ft2, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=201)
ap.add_argument("--nray", type=int, nargs="+", default=[251, 2501, 25001])
ap.add_argument("--scales", type=int, default=3)
args = ap.parse_args()
n, x, y = synth.shaped_density(args.nx, shift=0.0, elon=1.0)
a_set = np.linspace(0.2, 1.0, args.scales)
for _ in range(2):                                         # warm-up: CUDA context, library load, allocator
    ft2.phases(n, x, y, a_set=a_set, nray=251)
for nray in args.nray:
    row = {"nx": args.nx, "nray": nray, "beams": 7 * args.scales}
    for device in (True, False):
        if not device and nray > 3000:
            continue                                       # the host path is O(nray) Python objects
        best = None
        for _ in range(2 if nray <= 30000 else 1):
            t0 = time.perf_counter()
            ph_max, ph_sig = ft2.phases(n, x, y, a_set=a_set, nray=nray, device=device)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        row["device_s" if device else "host_s"] = round(best, 3)
        row["ph_sig_a1" if device else "ph_sig_a1_host"] = [round(float(v), 6) for v in ph_sig[-1, :3]]
    print(json.dumps(row), flush=True)
