"""-m gpu: the pooled handle lifecycle (vgs_acquire / vgs_release, what the drop-in classes use: one object per cloud,
test:51) and the staged pageable copies (csrc/vgs_hostio.cuh) give the oracle's results; getOneVoxelAdjacency through
vgs_get_unit_adjacency."""
import numpy as np
import pytest

from oracle import oracle
from vgs_svgs_segmentation_b200 import capi, scenes

from util import VGS_PARAMS

pytestmark = pytest.mark.gpu


def test_pooled_handle_is_clean_between_clouds(built_lib):
    """big cloud, then a small unrelated one on the SAME parked handle, then SVGS (other pool key), then the big one again"""
    L = capi.load()
    L.vgs_pool_trim()
    a = scenes.construction_site(300_000, seed=3, extent=13.0)
    b = scenes.two_planes(40_000, seed=9)
    ra, rb = oracle.run(a, math=1), oracle.run(b, math=1)
    first = None
    for pts, ref in ((a, ra), (b, rb), (a, ra)):
        h = capi.Handle(pooled=True)
        if first is None:
            first = h.h.value
        else:
            assert h.h.value == first, "the parked handle was not reused"
        h.set_points(pts)
        lab = h.run(capi.make_params(**VGS_PARAMS))
        assert h.counts()["n_points"] == pts.shape[0]
        np.testing.assert_array_equal(lab, ref.point_label)
        off, idx = h.clusters_csr(VGS_PARAMS["voxels_min"])
        assert len(off) - 1 == ref.stats["n_clusters_exported"]
        h.close()
    hs = capi.Handle(mode=capi.VGS_MODE_SVGS, pooled=True)      # another (device, mode) key: a different handle
    assert hs.h.value != first
    hs.close()
    L.vgs_pool_trim()


def test_staged_pageable_copies_equal_oracle(built_lib):
    """N large enough that the points (12 MB), labels and cluster indices (> 8 MB) go through the pinned staging ring"""
    pts = scenes.construction_site(3_000_000, seed=8, extent=38.0)
    import os
    oracle.set_threads(os.cpu_count() or 1)
    try:
        ref = oracle.run(pts, math=1)
    finally:
        oracle.set_threads(1)
    h = capi.Handle(pooled=True)
    h.set_points(pts)                      # numpy memory is pageable
    lab = h.run(capi.make_params(**VGS_PARAMS))
    np.testing.assert_array_equal(lab, ref.point_label)
    off, idx = h.clusters_csr(VGS_PARAMS["voxels_min"])
    assert idx.shape[0] > (8 << 20) // 4
    got = np.full(pts.shape[0], -1, np.int32)
    for c in range(len(off) - 1):
        m = idx[off[c]:off[c + 1]]
        got[m] = m.min()
    np.testing.assert_array_equal(got, ref.point_label)
    keys = h.blob("POINT_KEY").reshape(-1, 3)     # 26 MB blob through the same ring
    np.testing.assert_array_equal(keys, ref.point_key)
    h.close()


def test_unit_adjacency_slices(built_lib):
    pts = scenes.construction_site(120_000, seed=2, extent=8.0)
    ref = oracle.run(pts, math=1)
    h = capi.Handle()
    h.set_points(pts)
    h.voxelize(VGS_PARAMS["voxel_size"])
    h.compute_features(VGS_PARAMS["points_min"])
    h.find_adjacency(VGS_PARAMS["graph_size"])
    nv = h.voxel_count()
    for u in (0, 1, nv // 2, nv - 1):
        np.testing.assert_array_equal(h.unit_adjacency(u), ref.adj_idx[ref.adj_offsets[u]:ref.adj_offsets[u + 1]])
    with pytest.raises(capi.VgsError):
        h.unit_adjacency(nv)
    h.close()
