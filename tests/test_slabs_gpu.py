"""-m gpu: ONE scene split into spatial slabs (vgs_group_*) gives the labels of the single-device run
— and of the CPU oracle — bit for bit.  The loopback group runs all ranks of the split on one GPU with the schedule
the NCCL group uses (exchanges are device copies), so the split is covered on a single-GPU box; the NCCL transport
itself is covered when two devices are present."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import oracle
from vgs_svgs_segmentation_b200 import capi, scenes, slabs

from util import VGS_PARAMS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _single(xyz, leaf_order=0, **kw):
    p = dict(VGS_PARAMS); p.update(kw)
    h = capi.Handle(mode=0, leaf_order=leaf_order)
    h.set_points(xyz)
    lab = h.run(capi.make_params(**p))
    c = h.counts()
    h.close()
    return lab, c


def _split(xyz, nranks, leaf_order=0, bounds=None, **kw):
    p = dict(VGS_PARAMS); p.update(kw)
    n = xyz.shape[0]
    b = bounds or [slabs.slice_bounds(n, nranks, r) for r in range(nranks)]
    g = capi.Group(nranks, leaf_order=leaf_order)
    try:
        labs = g.run(capi.make_params(**p), [xyz[s:e] for s, e in b])
        return np.concatenate(labs), g.counts(), g.timings()
    finally:
        g.close()


@pytest.mark.parametrize("nranks", [1, 2, 3, 5, 8])
def test_slabs_equal_single_device_site(built_lib, nranks):
    xyz = scenes.construction_site(400_000, seed=3, extent=14.0)
    ref, c1 = _single(xyz)
    got, c, tm = _split(xyz, nranks)
    print(nranks, c, tm)
    np.testing.assert_array_equal(got, ref)
    assert c["n_clusters_exported"] == c1["n_clusters_exported"]
    assert c["n_points"] == xyz.shape[0]
    if nranks > 1:
        assert c["n_cross_pairs"] > 0, "the scene's ground plane crosses every cut: there must be merged components"


def test_slabs_equal_oracle_town(built_lib):
    xyz = scenes.town(200_000, seed=20170610, extent=11.0)
    r = oracle.run(xyz, math=1, **VGS_PARAMS)
    for nranks in (2, 4):
        got, c, _ = _split(xyz, nranks)
        np.testing.assert_array_equal(got, r.point_label)
        assert c["n_clusters_exported"] == r.stats["n_clusters_exported"]


def test_slabs_ascending_leaf_order_and_scan_order(built_lib):
    # insertion order drives PCL's bounding box: an unshuffled cloud grows the box many times, on every rank's slice
    xyz = scenes.construction_site(150_000, seed=5, extent=9.0, shuffle=False)
    for lo in (0, 1):
        ref, _ = _single(xyz, leaf_order=lo)
        got, c, _ = _split(xyz, 3, leaf_order=lo)
        np.testing.assert_array_equal(got, ref)
        assert c["origin_rounds"] >= 2


def test_slabs_nonfinite_points_uneven_and_empty_slices(built_lib):
    xyz = scenes.two_planes(60_000, seed=7).copy()
    xyz[5] = np.nan
    xyz[40_000, 1] = np.inf
    xyz[59_999] = np.nan
    ref, _ = _single(xyz)
    n = xyz.shape[0]
    got, c, _ = _split(xyz, 4, bounds=[(0, 10), (10, 10), (10, 45_000), (45_000, n)])   # rank 1 holds no points
    np.testing.assert_array_equal(got, ref)
    assert got[5] == -1 and got[40_000] == -1


def test_slabs_more_ranks_than_the_scene_is_wide(built_lib):
    # slabs thinner than the halo: a point is routed to several neighbours, some slabs own almost nothing
    xyz = scenes.two_planes(30_000, seed=9)
    ref, _ = _single(xyz)
    got, c, _ = _split(xyz, 8)
    np.testing.assert_array_equal(got, ref)


def test_slabs_count_slot_partner_far_away(built_lib):
    # closestCheck reads the neighbour COUNT as a voxel id (VS.h:2243): singles attach to one of the globally first
    # voxels wherever it lies; adjacency_min = 0 makes every single eligible so that the path is exercised a lot
    xyz = scenes.construction_site(200_000, seed=11, extent=12.0)
    for kw in (dict(adjacency_min=0), dict(adjacency_min=0, cut_thred=0.6), dict(points_min=2, adjacency_min=1)):
        ref, _ = _single(xyz, **kw)
        got, c, _ = _split(xyz, 4, **kw)
        np.testing.assert_array_equal(got, ref)


def test_slabs_run_twice_on_one_group(built_lib):
    a = scenes.construction_site(200_000, seed=3, extent=12.0)
    b = scenes.town(120_000, seed=4, extent=9.0)
    g = capi.Group(3)
    try:
        for xyz in (a, b, a):
            ref, _ = _single(xyz)
            bnd = [slabs.slice_bounds(xyz.shape[0], 3, r) for r in range(3)]
            got = np.concatenate(g.run(capi.make_params(**VGS_PARAMS), [xyz[s:e] for s, e in bnd]))
            np.testing.assert_array_equal(got, ref)
    finally:
        g.close()


NCCL_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    from vgs_svgs_segmentation_b200 import capi, scenes, slabs
    from util import VGS_PARAMS
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xyz = scenes.construction_site(600_000, seed=3, extent=16.0)
    s, e = slabs.slice_bounds(xyz.shape[0], world, rank)
    g = slabs.create_group(dist, rank, world, local)
    lab = g.run(capi.make_params(**VGS_PARAMS), [xyz[s:e]])[0]
    c = g.counts()
    g.close()
    h = capi.Handle(mode=0, device=local); h.set_points(xyz); ref = h.run(capi.make_params(**VGS_PARAMS)); h.close()
    assert np.array_equal(lab, ref[s:e]), (rank, int((lab != ref[s:e]).sum()))
    assert c["n_cross_pairs"] > 0
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_slabs_nccl_two_ranks(built_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the loopback tests cover the split on one)")
    script = tmp_path / "w.py"
    script.write_text(NCCL_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
