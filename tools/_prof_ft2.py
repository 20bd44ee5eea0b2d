import cProfile, pstats, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rayrtrace_b200 import ft2, synth
n, x, y = synth.shaped_density(201, shift=0.0, elon=1.0)
ft2.phases(n, x, y, a_set=[0.5], nray=51)
pr = cProfile.Profile(); pr.enable()
ft2.phases(n, x, y, a_set=[1.0], nray=25001)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
