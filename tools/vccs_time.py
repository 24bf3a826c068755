"""Time the supervoxel generator + SVGS on a large cloud (dev tool):  python tools/vccs_time.py [n_points]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vgs_svgs_segmentation_b200 import capi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = scenes.construction_site(n, seed=1, extent=70.0 * (n / 10_000_000) ** 0.5)
dev = torch.from_numpy(pts).cuda()
h = capi.Handle(mode=capi.VGS_MODE_SVGS, stream=torch.cuda.current_stream().cuda_stream)
p = capi.make_params(voxel_size=0.05, graph_size=0.5, sig_w=1.0, cut_thred=0.5)
for it in range(3):
    h.set_points_device(dev.data_ptr(), n, 12)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.voxelize(0.05)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    h.make_supervoxels_vccs(0.25, 0.0, 0.25, 0.75, 5)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    h.compute_features(10); h.find_adjacency(0.5); h.segment(p.sig, 0.5, 3)
    lab = h.point_labels(0)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    c = h.counts()
    print(f"n {n} voxels {h.voxel_count()} supervoxels {c['n_units']} clusters {c['n_clusters_all']}: voxelize {1e3*(t1-t0):.1f} ms, "
          f"vccs {1e3*(t2-t1):.1f} ms, svgs graph + labels {1e3*(t3-t2):.1f} ms")
