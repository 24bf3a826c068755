"""Stage-by-stage GPU-vs-oracle comparison that does not stop at the first mismatch (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
g.build(quiet=True)
from oracle import oracle
from vgs_svgs_segmentation_b200 import scenes, capi
from util import gpu_stages, csr_sets, oracle_conn_sets

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  [FAIL] {name}: shape {a.shape} vs {b.shape}"); return False
    if a.dtype.kind == 'f':
        bad = a.view(np.uint32) != b.view(np.uint32)
    else:
        bad = a != b
    nb = int(bad.sum())
    if nb:
        idx = np.argwhere(bad)[:5]
        print(f"  [FAIL] {name}: {nb}/{a.size} differ; first at {idx.tolist()} gpu={[a[tuple(i)] for i in idx]} ref={[b[tuple(i)] for i in idx]}")
        return False
    print(f"  [ok]   {name} ({a.size})"); return True

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "two_planes"
    npts = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    if which == "two_planes": xyz = scenes.two_planes(npts, seed=7)
    elif which == "site": xyz = scenes.construction_site(npts, seed=1, extent=12.0 * (npts / 250000) ** 0.5)
    else: xyz = scenes.town(npts, extent=11.0 * (npts / 200000) ** 0.5)
    t = time.time(); r = oracle.run(xyz, math=1); print("oracle s", time.time() - t, r.stats)
    t = time.time(); gg = gpu_stages(xyz); print("gpu s", time.time() - t, gg["counts"], gg["timings"])
    cmp("bbox", gg["bbox"], r.bbox)
    cmp("point_key", gg["point_key"], r.point_key)
    cmp("point_unit", gg["point_unit"], r.point_unit)
    cmp("unit_key", gg["unit_key"], r.unit_key)
    cmp("unit_center", gg["unit_center"], r.unit_center)
    cmp("unit_offsets", gg["unit_offsets"], r.unit_offsets)
    cmp("unit_points", gg["unit_points"], r.unit_points)
    cmp("used", gg["used"], r.used)
    cmp("centroid", gg["centroid"], r.centroid)
    cmp("normal", gg["normal"], r.normal)
    cmp("eigen", gg["eigen"], r.eigen)
    if cmp("adj_offsets", gg["adj_offsets"], r.adj_offsets):
        cmp("adj_idx", gg["adj_idx"], r.adj_idx)
        off = gg["adj_offsets"]
        for nm, ck, ik, ro, ri in (("conn0", "conn0_count", "conn0_idx", r.conn0_offsets, r.conn0_idx),
                                   ("conn1", "conn1_count", "conn1_idx", r.conn1_offsets, r.conn1_idx)):
            a = csr_sets(off, gg[ck], gg[ik]); b = oracle_conn_sets(ro, ri)
            bad = [i for i in range(len(a)) if a[i] != b[i]]
            print(f"  [{'ok' if not bad else 'FAIL'}]   {nm}: {len(bad)}/{len(a)} lists differ", [(i, a[i], b[i]) for i in bad[:3]])
    cmp("attach", gg["attach"], r.attach)
    cmp("point_label", gg["point_label"], r.point_label)
    print("clusters gpu", gg["n_clusters"], "oracle", r.stats["n_clusters_all"], r.stats["n_clusters_exported"])

main()
