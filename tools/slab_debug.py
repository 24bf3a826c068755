"""dev: where do the labels of a slab split differ from the single-device run?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from vgs_svgs_segmentation_b200 import capi, scenes, slabs
from util import VGS_PARAMS, gpu_stages

scene, nranks = sys.argv[1], int(sys.argv[2])
if scene == "town":
    xyz = scenes.town(200_000, seed=20170610, extent=11.0)
else:
    xyz = scenes.construction_site(400_000, seed=3, extent=14.0)
g1 = gpu_stages(xyz)
ref = g1["point_label"]
n = xyz.shape[0]
b = [slabs.slice_bounds(n, nranks, r) for r in range(nranks)]
g = capi.Group(nranks)
got = np.concatenate(g.run(capi.make_params(**VGS_PARAMS), [xyz[s:e] for s, e in b]))
c = g.counts()
print(c)
bad = np.nonzero(got != ref)[0]
print("mismatched points", bad.size)
pu = g1["point_unit"]
units = np.unique(pu[bad])
key = g1["unit_key"]; root = g1["unit_root"]; att = g1["attach"]; c1 = g1["conn1_count"]
off = g1["adj_offsets"]; adj = g1["adj_idx"]
ax = c["axis"]
for u in units[:40]:
    pts = bad[pu[bad] == u]
    r = root[u]
    members = np.nonzero(root == r)[0]
    print(f"unit {u} key {key[u]} axis-key {key[u][ax]} root {r} cluster-size {members.size} attach {att[u]} conn1 {c1[u]} nadj {off[u+1]-off[u]} "
          f"ref {ref[pts[0]]} got {got[pts[0]]} members-axis-range {key[members][:, ax].min()}..{key[members][:, ax].max()} "
          f"members {members[:8].tolist()} member-attach {att[members][:8].tolist()}")
