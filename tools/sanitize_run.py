"""Small VGS / SVGS / supervoxel-generator / partial-range runs for compute-sanitizer (dev tool):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py [--knobs]
The default paths take ~1 minute under memcheck; --knobs adds the alternative kernels (CTA local graph without pair
cache etc.), which take > 10 minutes under memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vgs_svgs_segmentation_b200 import capi, scenes
xyz = scenes.two_planes(30_000)
p = capi.make_params()
h = capi.Handle()
h.set_points(xyz)
lab = h.run(p)
print("vgs clusters", h.counts()["n_clusters_exported"])
# partial ranges + finish (multi-GPU path on one device)
h2 = capi.Handle()
h2.set_points(xyz)
h2.voxelize(p.voxel_size); h2.compute_features(p.points_min); h2.find_adjacency(p.graph_size)
nu = h2.unit_count()
h2.segment_partial(p.sig, p.cut_thred, 0, nu // 2)
h2.segment_partial(p.sig, p.cut_thred, nu // 2, nu)
h2.segment_finish(p.sig, p.cut_thred, p.adjacency_min)
assert np.array_equal(h2.point_labels(p.voxels_min), lab)
# SVGS with the built-in generator
ps = capi.make_params(voxel_size=0.05, graph_size=0.5, sig_w=1.0, cut_thred=0.5)
hs = capi.Handle(mode=capi.VGS_MODE_SVGS)
hs.set_points(xyz)
hs.voxelize(0.05)
hs.make_supervoxels_vccs()
hs.compute_features(10); hs.find_adjacency(0.5); hs.segment(ps.sig, 0.5, 3)
print("svgs clusters", hs.counts()["n_clusters_all"], "supervoxels", hs.counts()["n_units"])
hs.point_labels(0)
for knob in (("VGS_B200_NO_BITMAP", "VGS_B200_NO_WARP_KERNEL", "VGS_B200_NO_PAIR_CACHE", "VGS_B200_ADJ_TWO_PASS") if "--knobs" in sys.argv else ()):
    os.environ[knob] = "1"
    hk = capi.Handle()
    hk.set_points(xyz)
    assert np.array_equal(hk.run(p), lab), knob
    del os.environ[knob]
print("SANITIZE_RUN_OK")
