"""Regenerates tests/golden/*.npz from the CPU oracle (math=1, PCL-1.8.1 leaf order).

The reference has no golden vectors of its own and cannot be built here, so these fixtures freeze
the ORACLE's outputs on seeded synthetic scenes: they guard the oracle against regressions and are
a second, file-based target for the CUDA parity tests.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle  # noqa: E402
from vgs_svgs_segmentation_b200 import scenes  # noqa: E402

CASES = {
    "vgs_two_planes_20k": dict(scene=("two_planes", dict(n_points=20_000, seed=3)), params=dict(mode=0)),
    "vgs_site_40k": dict(scene=("construction_site", dict(n_points=40_000, seed=9, extent=5.0)), params=dict(mode=0)),
}


# supervoxel generator (oracle/vccs_oracle.cpp): Task_File_SVGS.txt values, both schedules
VCCS_CASES = {
    "vccs_two_planes_20k": dict(scene=("two_planes", dict(n_points=20_000, seed=3)),
                                params=dict(voxel_res=0.05, seed_res=0.25, color_importance=0.0, spatial_importance=0.25,
                                            normal_importance=0.75, refine_iterations=5)),
}


def make_scene(spec):
    name, kw = spec
    return getattr(scenes, name)(**kw)


def main():
    for name, case in CASES.items():
        xyz = make_scene(case["scene"])
        r = oracle.run(xyz, math=1, **case["params"])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            xyz_sha=np.frombuffer(__import__("hashlib").sha256(xyz.tobytes()).digest(), np.uint8),
            bbox=r.bbox, unit_key=r.unit_key, unit_offsets=r.unit_offsets, used=r.used,
            centroid=r.centroid, normal=r.normal, eigen=r.eigen,
            adj_offsets=r.adj_offsets, adj_idx=r.adj_idx,
            conn1_offsets=r.conn1_offsets, conn1_idx=r.conn1_idx, attach=r.attach,
            unit_cluster=r.unit_cluster, point_label=r.point_label,
            stats=np.array([r.stats[k] for k in oracle.STAT_NAMES], np.int64))
        print(name, r.stats)
    for name, case in VCCS_CASES.items():
        xyz = make_scene(case["scene"])
        seq = oracle.vccs(xyz, schedule=0, **case["params"])
        syn = oracle.vccs(xyz, schedule=1, **case["params"])
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            xyz_sha=np.frombuffer(__import__("hashlib").sha256(xyz.tobytes()).digest(), np.uint8),
            label_sequential=seq.point_label, max_label_sequential=np.int32(seq.max_label),
            label_synchronous=syn.point_label, max_label_synchronous=np.int32(syn.max_label),
            vox_normal=syn.vox_normal, n_seeds=np.int32(syn.n_seeds))
        print(name, "seeds", syn.n_seeds, "max label", seq.max_label, syn.max_label)


if __name__ == "__main__":
    main()
