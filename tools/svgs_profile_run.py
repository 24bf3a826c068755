import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from vgs_svgs_segmentation_b200 import capi
pts = bench.make_scene("town2m_svgs")
pd = bench.params_of("town2m_svgs")
h = capi.Handle(mode=1)
lab = np.empty(pts.shape[0], np.int32)
for _ in range(2):
    h.set_points(pts)
    bench.gpu_step(h, capi, "town2m_svgs", pd, lab, False)
print(h.counts())
for k in h.kernel_timings(): print(round(k["ms"], 3), k["launches"], k["name"])
