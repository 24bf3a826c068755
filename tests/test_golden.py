"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
not-gpu: the oracle still reproduces them; gpu: the CUDA path reproduces them through the C ABI."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from make_golden import CASES, VCCS_CASES, make_scene  # noqa: E402
from oracle import oracle  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    gold = _load(name)
    xyz = make_scene(CASES[name]["scene"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(xyz.tobytes()).digest(), np.uint8), gold["xyz_sha"]), \
        "the scene generator no longer produces the bytes the fixture was made from"
    r = oracle.run(xyz, math=1, **CASES[name]["params"])
    for k in ("bbox", "unit_key", "unit_offsets", "used", "adj_offsets", "adj_idx", "conn1_offsets", "conn1_idx",
              "attach", "unit_cluster", "point_label"):
        np.testing.assert_array_equal(r[k], gold[k], err_msg=k)
    for k in ("centroid", "normal", "eigen"):
        np.testing.assert_array_equal(r[k].view(np.uint32), gold[k].view(np.uint32), err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(built_lib, name):
    from util import gpu_stages
    gold = _load(name)
    xyz = make_scene(CASES[name]["scene"])
    g = gpu_stages(xyz)
    np.testing.assert_array_equal(g["bbox"], gold["bbox"])
    np.testing.assert_array_equal(g["unit_key"], gold["unit_key"])
    np.testing.assert_array_equal(g["unit_offsets"], gold["unit_offsets"])
    np.testing.assert_array_equal(g["used"], gold["used"])
    np.testing.assert_array_equal(g["adj_offsets"], gold["adj_offsets"])
    np.testing.assert_array_equal(g["adj_idx"], gold["adj_idx"])
    np.testing.assert_array_equal(g["attach"], gold["attach"])
    np.testing.assert_array_equal(g["point_label"], gold["point_label"])
    for k in ("centroid", "normal", "eigen"):
        rel = np.abs(g[k] - gold[k]) / np.maximum(np.abs(gold[k]), 1e-30)
        assert rel.max() <= 1e-5, k     # north_star tolerance for per-voxel features (fp32)


@pytest.mark.parametrize("name", sorted(VCCS_CASES))
def test_oracle_vccs_reproduces_golden(name):
    gold = _load(name)
    xyz = make_scene(VCCS_CASES[name]["scene"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(xyz.tobytes()).digest(), np.uint8), gold["xyz_sha"])
    for sched, key in ((0, "sequential"), (1, "synchronous")):
        r = oracle.vccs(xyz, schedule=sched, **VCCS_CASES[name]["params"])
        assert r.max_label == int(gold["max_label_" + key])
        np.testing.assert_array_equal(r.point_label, gold["label_" + key])
    np.testing.assert_array_equal(r.vox_normal.view(np.uint32), gold["vox_normal"].view(np.uint32))
    assert r.n_seeds == int(gold["n_seeds"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(VCCS_CASES))
def test_cuda_vccs_reproduces_golden(built_lib, name):
    from vgs_svgs_segmentation_b200 import capi
    gold = _load(name)
    xyz = make_scene(VCCS_CASES[name]["scene"])
    p = VCCS_CASES[name]["params"]
    h = capi.Handle(mode=capi.VGS_MODE_SVGS)
    h.set_points(xyz)
    h.voxelize(p["voxel_res"])
    h.make_supervoxels_vccs(p["seed_res"], p["color_importance"], p["spatial_importance"], p["normal_importance"], p["refine_iterations"])
    lab, ml = h.supervoxel_labels()
    assert ml == int(gold["max_label_synchronous"])
    np.testing.assert_array_equal(lab, gold["label_synchronous"])
