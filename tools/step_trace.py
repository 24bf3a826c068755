"""Per-step timing trace (dev tool): python tools/step_trace.py [points] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build(oracle=False, quiet=True)
from vgs_svgs_segmentation_b200 import capi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pts = scenes.construction_site(n, seed=1, extent=70.0 * (n / 10_000_000) ** 0.5)
dev = torch.from_numpy(pts).cuda()
lab = torch.empty(n, dtype=torch.int32, device="cuda")
h = capi.Handle(stream=torch.cuda.current_stream().cuda_stream)
p = capi.make_params()
out = []
for i in range(steps):
    h.set_points_device(dev.data_ptr(), n, 12)
    t0 = time.perf_counter()
    h.run(p, lab.data_ptr(), on_device=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    tm = h.timings()
    out.append((round(dt, 1), round(tm["graph_ms"], 1), round(tm["pair_cache_ms"], 1)))
print(out)
