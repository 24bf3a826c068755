"""In-process A/B timing of tuning knobs (dev tool): every configuration gets its own handle (the env
knobs are read at vgs_create), steps are interleaved, medians are reported.
   python tools/ab_tune.py VGS_B200_LW_CHUNK 12 20 28 40 64"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build(oracle=False, quiet=True)
from vgs_svgs_segmentation_b200 import capi, scenes
knob, values = sys.argv[1], sys.argv[2:]
n = 10_000_000
pts = scenes.construction_site(n, seed=1, extent=70.0)
dev = torch.from_numpy(pts).cuda()
lab = torch.empty(n, dtype=torch.int32, device="cuda")
p = capi.make_params()
hs = {}
for v in values:
    os.environ[knob] = v
    hs[v] = capi.Handle(stream=torch.cuda.current_stream().cuda_stream)
res = {v: [] for v in values}
for it in range(9):
    for v in values:
        h = hs[v]
        h.set_points_device(dev.data_ptr(), n, 12)
        h.run(p, lab.data_ptr(), on_device=True)
        torch.cuda.synchronize()
        if it >= 2:
            res[v].append(h.timings()[os.environ.get("AB_KEY", "graph_ms")])
for v in values:
    print(knob, v, os.environ.get("AB_KEY", "graph_ms"), "median", round(statistics.median(res[v]), 2), "min", round(min(res[v]), 2), "max", round(max(res[v]), 2))
