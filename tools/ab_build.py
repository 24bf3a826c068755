"""In-process A/B timing of two builds of the library (dev tool): steps are interleaved on one device so
that box-to-box and process-to-process variance cancels.
   make -C vgs_svgs_segmentation_b200/csrc alt          # builds _alt/libvgs_b200_alt.so with -DVGS_AB_ALT
   python tools/ab_build.py vgs_svgs_segmentation_b200/libvgs_b200.so _alt/libvgs_b200_alt.so"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vgs_svgs_segmentation_b200 import capi, scenes
paths = sys.argv[1:]
n = 10_000_000
pts = scenes.construction_site(n, seed=1, extent=70.0)
dev = torch.from_numpy(pts).cuda()
labs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in paths]
p = capi.make_params()
hs = []
for path in paths:
    capi._lib, capi.SO_PATH = None, os.path.abspath(path)
    hs.append(capi.Handle(stream=torch.cuda.current_stream().cuda_stream))
keys = ("total_ms", "graph_ms", "pair_cache_ms", "adjacency_ms", "mutual_ms", "components_ms", "labels_ms", "voxelize_ms")
res = [{k: [] for k in keys} for _ in paths]
kres = [{} for _ in paths]
for it in range(13):
    for i, h in enumerate(hs):
        h.set_points_device(dev.data_ptr(), n, 12)
        h.run(p, labs[i].data_ptr(), on_device=True)
        torch.cuda.synchronize()
        if it >= 3:
            t = h.timings()
            for k in keys:
                res[i][k].append(t[k])
            for kt in h.kernel_timings():
                kres[i].setdefault(kt["name"].split(":")[1].strip()[:28], []).append(kt["ms"])
for i, path in enumerate(paths):
    print(os.path.basename(path), {k: (round(statistics.median(v), 2), round(min(v), 2)) for k, v in res[i].items()})
for i, path in enumerate(paths):
    print(os.path.basename(path), "kernels", {k: round(statistics.median(v), 3) for k, v in kres[i].items()})
print("labels equal:", all(bool((labs[0] == l).all()) for l in labs[1:]))
