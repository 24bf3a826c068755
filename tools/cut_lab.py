"""Drives tools/cut_lab.cpp: counts visited / useful entries of the local-graph cut per termination rule (design aid)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle
from vgs_svgs_segmentation_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
scene = sys.argv[2] if len(sys.argv) > 2 else "site"
if scene == "site":
    pts = scenes.construction_site(n, seed=1, extent=70.0 * (n / 1e7) ** 0.5)
else:
    pts = scenes.town(n, extent=60.0 * (n / 2e6) ** 0.5)
oracle.set_threads(8)
t0 = time.time(); r = oracle.run(pts, math=1); print("oracle", time.time() - t0, r.stats)
V = r.stats["n_units"]
rec = np.zeros((V, 16), np.float32)
rec[:, 0:3] = r.centroid; rec[:, 3:6] = r.normal; rec[:, 6:14] = r.eigen
cnt = np.diff(r.unit_offsets).astype(np.int32)
rec[:, 14] = cnt.view(np.float32)
used = r.used.astype(bool)
fl = np.zeros(V, np.int32)
fl |= np.where((rec[:, 0] != 0) & (rec[:, 1] != 0) & (rec[:, 2] != 0), 1, 0)
fl |= np.where((rec[:, 3] != 0) & (rec[:, 4] != 0) & (rec[:, 5] != 0), 2, 0)
fl |= np.where(used, 12, 0)
fl = np.where(used, fl, 0).astype(np.int32)
rec[:, 15] = fl.view(np.float32)
L = C.CDLL(os.path.join(ROOT, "tests/_build/libcut_lab.so"))
stats = np.zeros(16, np.int64); cc = np.zeros(V, np.int32); ci = np.zeros(len(r.adj_idx), np.int32); hist = np.zeros(101, np.int32)
sig = np.array([0.2, 0.2, 0.2, 0.2, 0.2, 2.0], np.float32)
ao = np.ascontiguousarray(r.adj_offsets, np.int64); ai = np.ascontiguousarray(r.adj_idx, np.int32)
t0 = time.time()
L.lab_run(rec.ctypes.data_as(C.c_void_p), ao.ctypes.data_as(C.c_void_p), ai.ctypes.data_as(C.c_void_p), C.c_int64(V), sig.ctypes.data_as(C.c_void_p),
          C.c_float(0.3), stats.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), hist.ctypes.data_as(C.c_void_p))
print("lab", time.time() - t0)
names = ["units", "sum_nv", "sum_pairs", "kept", "vis_min", "vis_s0", "use_min", "use_s0", "single_shortcut", "end_nseg1", "sum_S0", "merges"]
u = stats[0]
for k, v in zip(names, stats):
    print(f"{k:16s} {v:12d}  per unit {v / u:9.2f}")
print("mismatch", stats[15])
# compare with the oracle's conn0
ok = 0; bad = 0
c0o = r.conn0_offsets
for v in range(V):
    a = np.sort(r.conn0_idx[c0o[v]:c0o[v + 1]]); b = np.sort(ci[ao[v]:ao[v] + cc[v]])
    if len(a) == len(b) and (a == b).all(): ok += 1
    else: bad += 1
print("conn0 equal", ok, "different", bad)
print("visited% histogram (S0 rule):", hist.tolist())
