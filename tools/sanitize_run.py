"""Small VGS / SVGS / supervoxel-generator / slab-group runs for compute-sanitizer (dev tool):
   compute-sanitizer --tool memcheck  python tools/sanitize_run.py [--knobs]
   compute-sanitizer --tool racecheck python tools/sanitize_run.py
The default paths take a few minutes under memcheck; --knobs adds the alternative kernels (general local-graph kernel
without weight rows, forced fallback units, hash lookups without the id grid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vgs_svgs_segmentation_b200 import capi, scenes, slabs
xyz = scenes.two_planes(30_000)
p = capi.make_params()
h = capi.Handle()
h.set_points(xyz)
lab = h.run(p)
off, idx = h.clusters_csr(p.voxels_min)
print("vgs clusters", h.counts()["n_clusters_exported"], "csr", len(off) - 1, idx.shape[0])
# one scene split over 3 ranks on this device (the schedule of the NCCL group, exchanges by device copies)
g = capi.Group(3)
b = [slabs.slice_bounds(xyz.shape[0], 3, r) for r in range(3)]
got = np.concatenate(g.run(p, [xyz[s:e] for s, e in b]))
assert np.array_equal(got, lab), "slab group differs"
g.close()
# SVGS with the built-in generator
ps = capi.make_params(voxel_size=0.05, graph_size=0.5, sig_w=1.0, cut_thred=0.5)
hs = capi.Handle(mode=capi.VGS_MODE_SVGS)
hs.set_points(xyz)
hs.voxelize(0.05)
hs.make_supervoxels_vccs()
hs.compute_features(10); hs.find_adjacency(0.5); hs.segment(ps.sig, 0.5, 3)
print("svgs clusters", hs.counts()["n_clusters_all"], "supervoxels", hs.counts()["n_units"])
hs.point_labels(0)
for knob, val in ((("VGS_B200_NO_PAIR_CACHE", "1"), ("VGS_B200_FORCE_FALLBACK", "3"), ("VGS_B200_IDGRID_MB", "0"), ("VGS_B200_LG_ORDER", "8"))
                  if "--knobs" in sys.argv else ()):
    os.environ[knob] = val
    hk = capi.Handle()
    hk.set_points(xyz)
    assert np.array_equal(hk.run(p), lab), knob
    del os.environ[knob]
print("SANITIZE_RUN_OK")
