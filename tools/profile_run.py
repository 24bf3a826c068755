"""One pipeline pass for profiling under ncu (no warm-up loop): python tools/profile_run.py [points] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build(oracle=False, quiet=True)
from vgs_svgs_segmentation_b200 import capi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
extent = 70.0 * (n / 10_000_000) ** 0.5
pts = scenes.construction_site(n, seed=1, extent=extent)
h = capi.Handle()
for _ in range(passes):
    h.set_points(pts)
    lab = h.run(capi.make_params())
print(h.counts(), h.timings())
