#!/usr/bin/env python
"""cProfile of one application-level call (ft2.phases shape: one density, 7 beams x nray rays) on the GPU box:
where the wall time of the O(rays) post-processing goes."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayrtrace_b200 import ft2, synth  # noqa: E402

nray = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_003
n, x, y = synth.shaped_density(201, shift=0.0, elon=1.0)
ft2.phases(n, x, y, a_set=[1.0], nray=251)
ft2.phases(n, x, y, a_set=[1.0], nray=20001)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
ph_max, ph_sig = ft2.phases(n, x, y, a_set=[1.0], nray=nray)
pr.disable()
print(f"wall {time.perf_counter() - t0:.3f} s for 7 x {nray} rays; ph_sig {np.round(ph_sig[-1, :3], 6)}")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
