"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, stage by stage.

Bar (BASELINE.json north_star): voxel keys, adjacency lists and canonical labels bit-exact;
per-unit features within 1e-5 relative (they are in fact compared bit for bit first and the number
of non-identical floats is reported)."""
import numpy as np
import pytest

from oracle import oracle
from vgs_svgs_segmentation_b200 import scenes

from util import VGS_PARAMS, csr_sets, gpu_stages, oracle_conn_sets


def capi_handle(**kw):
    from vgs_svgs_segmentation_b200 import capi
    return capi.Handle(**kw)

pytestmark = pytest.mark.gpu

FEATURE_RTOL = 1e-5   # north_star: "per-voxel features agree within 1e-5 relative in fp32"


def _scene(name):
    if name == "two_planes":
        return scenes.two_planes(60_000, seed=7)
    if name == "site":
        return scenes.construction_site(250_000, seed=1, extent=12.0)
    if name == "town":
        return scenes.town(200_000, seed=20170610, extent=11.0)
    if name == "site_scanorder":
        return scenes.construction_site(120_000, seed=5, extent=8.0, shuffle=False)
    raise KeyError(name)


def _compare_vgs(xyz, g, r, check_labels=True):
    np.testing.assert_array_equal(g["bbox"], r.bbox)                       # PCL dynamic bounding box
    assert g["n_voxels"] == r.stats["n_units"]
    np.testing.assert_array_equal(g["point_key"], r.point_key)             # octree keys, bit-exact
    np.testing.assert_array_equal(g["point_unit"], r.point_unit)           # voxel ids = leaf order
    np.testing.assert_array_equal(g["unit_key"], r.unit_key)
    np.testing.assert_array_equal(g["unit_center"].view(np.uint32), r.unit_center.view(np.uint32))
    np.testing.assert_array_equal(g["unit_offsets"], r.unit_offsets)
    np.testing.assert_array_equal(g["unit_points"], r.unit_points)         # ascending indices per voxel
    np.testing.assert_array_equal(g["used"], r.used)
    for k in ("centroid", "normal", "eigen"):
        a, b = g[k], r[k]
        nbad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        denom = np.maximum(np.abs(b), 1e-30)
        rel = float((np.abs(a - b) / denom).max()) if a.size else 0.0
        print(f"{k}: {nbad} of {a.size} floats not bit-identical, max rel err {rel:.3g}")
        assert rel <= FEATURE_RTOL, k
        assert nbad == 0, f"{k}: expected bit-identical features (sequential fp32 sums + correctly rounded libm)"
    np.testing.assert_array_equal(g["adj_offsets"], r.adj_offsets)
    np.testing.assert_array_equal(g["adj_idx"], r.adj_idx)                 # ordered (dist2, id) lists
    off = g["adj_offsets"]
    assert csr_sets(off, g["conn0_count"], g["conn0_idx"]) == oracle_conn_sets(r.conn0_offsets, r.conn0_idx)
    assert csr_sets(off, g["conn1_count"], g["conn1_idx"]) == oracle_conn_sets(r.conn1_offsets, r.conn1_idx)
    np.testing.assert_array_equal(g["attach"], r.attach)
    np.testing.assert_array_equal(g["unit_root"], _roots(r.unit_cluster))
    assert g["n_clusters"] == (r.stats["n_clusters_all"], r.stats["n_clusters_exported"])
    if check_labels:
        np.testing.assert_array_equal(g["point_label"], r.point_label)     # canonical labels
    # exported clusters: same cluster order, same point sets
    goff, gidx = g["clusters_csr"]
    assert len(goff) == len(r.cluster_offsets)
    np.testing.assert_array_equal(goff, r.cluster_offsets)
    for c in range(len(goff) - 1):
        assert set(gidx[goff[c]:goff[c + 1]].tolist()) == set(r.cluster_points[r.cluster_offsets[c]:r.cluster_offsets[c + 1]].tolist())


def _roots(unit_cluster):
    """oracle cluster index -> smallest unit id of that cluster (the CUDA path's root)"""
    first = {}
    for u, c in enumerate(unit_cluster.tolist()):
        first.setdefault(c, u)
    return np.array([first[c] for c in unit_cluster.tolist()], dtype=np.int32)


@pytest.mark.parametrize("name", ["two_planes", "site", "town", "site_scanorder"])
def test_vgs_stage_parity(built_lib, name):
    xyz = _scene(name)
    g = gpu_stages(xyz)
    r = oracle.run(xyz, math=1)
    print(name, g["counts"], g["timings"])
    _compare_vgs(xyz, g, r)


def test_vgs_ascending_leaf_order(built_lib):
    xyz = _scene("two_planes")
    g = gpu_stages(xyz, leaf_order=1)
    r = oracle.run(xyz, math=1, leaf_order=1)
    _compare_vgs(xyz, g, r)


def test_vgs_other_parameters(built_lib):
    """coarser voxels, smaller graph radius (27-stencil regime), looser cut, no min sizes"""
    xyz = _scene("site")
    kw = dict(voxel_size=0.2, graph_size=0.39, cut_thred=0.45, points_min=4, adjacency_min=1, voxels_min=0, sig_w=1.5)
    g = gpu_stages(xyz, **kw)
    r = oracle.run(xyz, math=1, **kw)
    _compare_vgs(xyz, g, r)


def test_vgs_stride16_and_nonfinite(built_lib):
    xyz = _scene("two_planes")
    xyz4 = np.zeros((xyz.shape[0], 4), np.float32)
    xyz4[:, :3] = xyz
    xyz4[:, 3] = 1.0
    xyz4[5, 0] = np.nan
    xyz4[777, 2] = np.inf
    xyz4[0, 1] = np.nan          # the very first point is skipped: the box starts at point 1
    g = gpu_stages(xyz4)
    r = oracle.run(xyz4, math=1)
    assert r.stats["n_finite"] == xyz.shape[0] - 3
    _compare_vgs(xyz4, g, r)
    assert g["point_label"][5] == -1 and g["point_label"][777] == -1


def test_vgs_glibc_libm_differences_are_near_threshold(built_lib):
    """The literal GCC/glibc build of the reference (oracle math=0) differs from the correctly
    rounded definition only through float-libm rounding; any label difference must be explained by
    merge decisions within the reported near-threshold set."""
    xyz = _scene("site")
    g = gpu_stages(xyz)
    r0 = oracle.run(xyz, math=0, near_tol=1e-5)
    ndiff = int((g["point_label"] != r0.point_label).sum())
    print("label differences vs glibc-libm oracle:", ndiff, "near-threshold decisions:", r0.stats["near_threshold"])
    if ndiff:
        assert r0.stats["near_threshold"] > 0


def test_idempotent_rerun_and_run_api(built_lib):
    """vgs_run twice on one handle gives identical labels (buffers are reused, no stale state)."""
    from vgs_svgs_segmentation_b200 import capi
    xyz = _scene("two_planes")
    h = capi.Handle()
    h.set_points(xyz)
    a = h.run(capi.make_params(**VGS_PARAMS)).copy()
    h.set_points(xyz)
    b = h.run(capi.make_params(**VGS_PARAMS)).copy()
    h.close()
    np.testing.assert_array_equal(a, b)
    r = oracle.run(xyz, math=1)
    np.testing.assert_array_equal(a, r.point_label)


def test_state_errors(built_lib):
    from vgs_svgs_segmentation_b200 import capi
    h = capi.Handle()
    with pytest.raises(capi.VgsError) as e:
        h.voxelize(0.15)
    assert e.value.status == 3
    xyz = _scene("two_planes")
    h.set_points(xyz)
    with pytest.raises(capi.VgsError):
        h.find_adjacency(0.5)
    h.close()


def test_vgs_without_pair_cache(built_lib, monkeypatch):
    """VGS_B200_NO_PAIR_CACHE=1: weights evaluated inside every local graph by the general kernel (the SVGS code
    path) must give the same lists and labels as the weight rows."""
    monkeypatch.setenv("VGS_B200_NO_PAIR_CACHE", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    assert g["timings"]["pair_cache_ms"] == 0.0
    monkeypatch.delenv("VGS_B200_NO_PAIR_CACHE")
    g2 = gpu_stages(xyz)
    assert g2["timings"]["pair_cache_ms"] > 0.0
    r = oracle.run(xyz, math=1)
    _compare_vgs(xyz, g, r)
    _compare_vgs(xyz, g2, r)


def test_vgs_row_kernel_fallback_units(built_lib, monkeypatch):
    """VGS_B200_FORCE_FALLBACK=3: the row kernel hands every third voxel to the general kernel (the path taken by a voxel
    whose weights overflow the staging buffer inside one weight cell); their connect lists come back as lists and are
    converted to lattice masks — results must not change."""
    monkeypatch.setenv("VGS_B200_FORCE_FALLBACK", "3")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_FORCE_FALLBACK")
    assert g["counts"]["n_units"] > 0
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_hash_lookups_without_id_grid(built_lib, monkeypatch):
    """VGS_B200_IDGRID_MB=0: neighbour ids through the Morton hash table (the path of scenes whose occupied key range is too
    large for a dense id grid) — adjacency lists, weight rows and labels must not change."""
    monkeypatch.setenv("VGS_B200_IDGRID_MB", "0")
    xyz = _scene("town")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_IDGRID_MB")
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_single_stream_classes(built_lib, monkeypatch):
    """VGS_B200_CLASS_STREAMS=1 with the general kernel: all size-class launches on the handle's stream (default: 6
    streams, forked / joined with events); the connect lists do not depend on how the launches overlap."""
    monkeypatch.setenv("VGS_B200_CLASS_STREAMS", "1")
    monkeypatch.setenv("VGS_B200_NO_PAIR_CACHE", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_CLASS_STREAMS")
    monkeypatch.delenv("VGS_B200_NO_PAIR_CACHE")
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_dense_blob_long_rows(built_lib):
    """a solid random blob: neighbourhoods of 100+ voxels, weight rows longer than the 512-entry short-row sort (the
    long-row launch), staged rounds larger than one warp"""
    rng = np.random.default_rng(12)
    xyz = (rng.random((400_000, 3)) * np.array([2.4, 2.4, 2.4]) + 1.0).astype(np.float32)
    g = gpu_stages(xyz)
    assert g["counts"]["max_neighbours"] > 120
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_large_cut_keeps_zero_weights(built_lib):
    """cut_thred > 0.5 makes the cut bound negative: zero-weight entries (pairs with unused voxels)
    stay in play and may merge (SURVEY.md A.5 last bullet)."""
    xyz = _scene("two_planes")
    kw = dict(cut_thred=0.8)
    g = gpu_stages(xyz, **kw)
    r = oracle.run(xyz, math=1, **kw)
    _compare_vgs(xyz, g, r)


def _compare_svgs(xyz, labels, g, r):
    np.testing.assert_array_equal(g["point_unit"], r.point_unit)           # supervoxel ids = ascending label
    np.testing.assert_array_equal(g["unit_offsets"], r.unit_offsets)
    np.testing.assert_array_equal(g["unit_points"], r.unit_points)
    for k in ("centroid", "normal", "eigen"):
        a, b = g[k], r[k]
        nbad = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        rel = float((np.abs(a - b) / np.maximum(np.abs(b), 1e-30)).max())
        print(f"{k}: {nbad} of {a.size} floats not bit-identical, max rel err {rel:.3g}")
        assert rel <= FEATURE_RTOL and nbad == 0, k
    np.testing.assert_array_equal(g["adj_offsets"], r.adj_offsets)
    np.testing.assert_array_equal(g["adj_idx"], r.adj_idx)
    off = g["adj_offsets"]
    assert csr_sets(off, g["conn0_count"], g["conn0_idx"]) == oracle_conn_sets(r.conn0_offsets, r.conn0_idx)
    assert csr_sets(off, g["conn1_count"], g["conn1_idx"]) == oracle_conn_sets(r.conn1_offsets, r.conn1_idx)
    np.testing.assert_array_equal(g["attach"], r.attach)
    np.testing.assert_array_equal(g["unit_root"], _roots(r.unit_cluster))
    assert g["n_clusters"] == (r.stats["n_clusters_all"], r.stats["n_clusters_exported"])
    np.testing.assert_array_equal(g["point_label"], r.point_label)


@pytest.mark.parametrize("name,seed_size", [("town", 0.25), ("site", 0.25), ("two_planes", 0.2)])
def test_svgs_stage_parity(built_lib, name, seed_size):
    """SVGS downstream of supplied per-point supervoxel labels (the VCCS generator itself is PCL
    code with unpinned parity): Task_File_SVGS.txt parameters."""
    xyz = _scene(name)
    labels = scenes.supervoxel_labels_grid(xyz, seed_size)
    ml = int(labels.max()) + 1
    g = gpu_stages(xyz, mode=1, labels=labels, max_label=ml)
    r = oracle.run(xyz, labels=labels, max_label=ml, mode=1, math=1)
    print(name, g["counts"], g["timings"])
    assert g["n_voxels"] == r.stats["max_n_or_voxels"]          # getVoxelNum of the 0.05 m octree (SV.h:111)
    _compare_svgs(xyz, labels, g, r)


def test_svgs_drops_max_label_and_unlabelled(built_lib):
    """SV.h:303-323: label 0 is unlabelled, and the label loop stops before max_label."""
    xyz = _scene("two_planes")
    labels = scenes.supervoxel_labels_grid(xyz, 0.25)
    labels[::17] = 0
    ml = int(labels.max())          # getMaxLabel(): the largest label is dropped by `k < max_label`
    g = gpu_stages(xyz, mode=1, labels=labels, max_label=ml)
    r = oracle.run(xyz, labels=labels, max_label=ml, mode=1, math=1)
    _compare_svgs(xyz, labels, g, r)
    assert np.all(g["point_label"][labels == 0] == -1) and np.all(g["point_label"][labels == ml] == -1)


def _rows_sorted(offsets, counts, idx):
    """(row, id) pairs of CSR lists with `counts` valid entries per row, sorted by (row, id) — set comparison of lists at
    bench sizes without Python loops"""
    offsets = np.asarray(offsets, np.int64)
    counts = np.asarray(counts, np.int64)
    row = np.repeat(np.arange(len(counts), dtype=np.int64), counts)
    start = np.repeat(offsets[:-1], counts)
    within = np.arange(int(counts.sum()), dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
    ids = np.asarray(idx)[start + within].astype(np.int64)
    order = np.lexsort((ids, row))
    return row[order], ids[order]


def _full_size_vs_oracle(xyz, mode=0, labels=None, max_label=0, vccs=False):
    """one full run through the C ABI at a benchmarked size against the threaded oracle (math=1): canonical labels,
    adjacency lists (order included), connect lists after the mutual filter, closest-check partners, cluster roots"""
    from vgs_svgs_segmentation_b200 import capi
    from util import SVGS_PARAMS
    pd = dict(SVGS_PARAMS if mode == 1 else VGS_PARAMS)
    h = capi.Handle(mode=mode)
    h.set_points(xyz)
    h.voxelize(pd["voxel_size"])
    if mode == 1:
        if vccs:
            h.make_supervoxels_vccs(0.25, 0.0, 0.25, 0.75, 5)
            labels, max_label = h.supervoxel_labels()
        else:
            h.set_supervoxel_labels(labels, max_label)
    h.compute_features(pd["points_min"])
    h.find_adjacency(pd["graph_size"])
    h.segment(capi.Sigmas(pd["sig_p"], pd["sig_n"], pd["sig_o"], pd["sig_e"], pd["sig_c"], pd["sig_w"]), pd["cut_thred"], pd["adjacency_min"])
    lab = h.point_labels(pd["voxels_min"])
    g = dict(adj_offsets=h.blob("ADJ_OFFSETS"), adj_idx=h.blob("ADJ_IDX"), conn1_count=h.blob("CONN1_COUNT"), conn1_idx=h.blob("CONN1_IDX"),
             attach=h.blob("ATTACH"), root=h.blob("UNIT_ROOT"), counts=h.counts())
    h.close()
    import os
    oracle.set_threads(os.cpu_count() or 1)   # all host threads for the per-unit loop; results do not depend on it
    try:
        r = oracle.run(xyz, labels=labels, max_label=max_label, mode=mode, math=1, **{k: v for k, v in pd.items()})
    finally:
        oracle.set_threads(1)
    assert g["counts"]["n_units"] == r.stats["n_units"]
    np.testing.assert_array_equal(g["adj_offsets"], r.adj_offsets)
    np.testing.assert_array_equal(g["adj_idx"], r.adj_idx)
    grow, gid = _rows_sorted(g["adj_offsets"], g["conn1_count"], g["conn1_idx"])
    orow, oid = _rows_sorted(r.conn1_offsets, np.diff(r.conn1_offsets), r.conn1_idx)
    np.testing.assert_array_equal(grow, orow)
    np.testing.assert_array_equal(gid, oid)
    np.testing.assert_array_equal(g["attach"], r.attach)
    np.testing.assert_array_equal(lab, r.point_label)
    print(f"N={len(xyz)} units={r.stats['n_units']} used={r.stats['n_used']} clusters={r.stats['n_clusters_exported']} "
          f"near_threshold={r.stats['near_threshold']} labels equal: {len(lab)}/{len(lab)}")
    return lab, r


def test_bench_scene_10m_equals_oracle(built_lib):
    """BASELINE.json configs[2] at FULL size — the scene bench.py times (10 M points, seed 1, 70 m): labels, adjacency,
    connect lists and closest-check partners equal to the oracle; a second run is bit-identical"""
    from vgs_svgs_segmentation_b200 import capi
    xyz = scenes.construction_site(10_000_000, seed=1, extent=70.0)
    lab, r = _full_size_vs_oracle(xyz)
    h = capi.Handle()
    h.set_points(xyz)
    lab2 = h.run(capi.make_params(**VGS_PARAMS))
    h.close()
    assert np.array_equal(lab, lab2)


def test_town_2m_vgs_equals_oracle(built_lib):
    """configs[0] stand-in (Town_Test.pcd is not distributed): the 2 M-point town scene with Task_File_VGS.txt parameters"""
    _full_size_vs_oracle(scenes.town(2_000_000))


def test_town_2m_svgs_equals_oracle(built_lib):
    """configs[1] stand-in: SVGS with Task_File_SVGS.txt parameters on the 2 M-point town scene, supervoxels supplied
    (seed-grid labels, 0.25 m) and made by the CUDA VCCS generator (the oracle is fed the generator's labels)"""
    xyz = scenes.town(2_000_000)
    labels = scenes.supervoxel_labels_grid(xyz, 0.25)
    _full_size_vs_oracle(xyz, mode=1, labels=labels, max_label=int(labels.max()) + 1)
    _full_size_vs_oracle(xyz, mode=1, vccs=True)


def _tiny_clouds():
    rng = np.random.default_rng(5)
    one = np.array([[0.3, 0.4, 0.5]], np.float32)
    few = rng.normal(size=(5, 3)).astype(np.float32)
    same = np.repeat(np.array([[1.25, -2.5, 0.75]], np.float32), 40, axis=0)           # one voxel, zero scatter
    one_voxel = (np.array([[2.01, 3.01, 1.01]]) + 0.1 * rng.random((200, 3))).astype(np.float32)
    sparse = (rng.random((300, 3)) * 30).astype(np.float32)                               # every voxel unused
    line = np.stack([np.linspace(0, 3, 4000), np.full(4000, 0.7), np.full(4000, 0.2)], 1).astype(np.float32)
    line += 0.002 * rng.standard_normal(line.shape).astype(np.float32)
    return dict(one=one, few=few, same=same, one_voxel=one_voxel, sparse=sparse, line=line)


@pytest.mark.parametrize("name", ["one", "few", "same", "one_voxel", "sparse", "line"])
def test_vgs_tiny_and_degenerate_clouds(built_lib, name):
    """ragged / degenerate inputs: a single point, a handful, identical points (zero scatter matrix ->
    identity eigenvectors -> 'empty' normal), one voxel, all voxels unused, a 1-D line of voxels"""
    xyz = _tiny_clouds()[name]
    g = gpu_stages(xyz)
    r = oracle.run(xyz, math=1)
    _compare_vgs(xyz, g, r)


def test_vgs_far_from_origin(built_lib):
    """coordinates of ~5 km: the cross-product cue <X1 x X2, d> leaves [-1,1], acos gives NaN and the
    convexity cue falls back to PI; keys need 16+ bits per axis"""
    xyz = scenes.two_planes(40_000, seed=9) + np.array([5000.0, -3000.0, 150.0], np.float32)
    g = gpu_stages(xyz)
    r = oracle.run(xyz, math=1)
    _compare_vgs(xyz, g, r)


def test_vgs_deep_octree_two_sites(built_lib):
    """two patches 700 m apart: a 13-level octree (> 32-bit sort keys, occupancy / id grids over a 4 700-cell range) — the
    lattice searches must stay on the grid path and equal the oracle"""
    a = scenes.construction_site(120_000, seed=1, extent=8.0)
    b = scenes.construction_site(120_000, seed=2, extent=8.0) + np.array([700.0, 0.0, 0.0], np.float32)
    xyz = np.concatenate([a, b], axis=0)
    xyz = np.ascontiguousarray(xyz[np.random.default_rng(5).permutation(xyz.shape[0])])
    g = gpu_stages(xyz)
    assert g["counts"]["octree_depth"] >= 13
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_svgs_tiny(built_lib):
    rng = np.random.default_rng(2)
    xyz = (rng.random((500, 3)) * np.array([2.0, 2.0, 0.02])).astype(np.float32) + 0.5
    labels = scenes.supervoxel_labels_grid(xyz, 0.25)
    ml = int(labels.max()) + 1
    g = gpu_stages(xyz, mode=1, labels=labels, max_label=ml)
    r = oracle.run(xyz, labels=labels, max_label=ml, mode=1, math=1)
    _compare_svgs(xyz, labels, g, r)


def test_origin_growth_loop_late_violations(built_lib):
    """the device-side growth loop of PCL's bounding box (k_origin_scan / k_origin_adopt): violations far apart in the
    insertion order make the scan windows miss and grow (x64), and the ten rounds of the first batch are not enough —
    box, depth and every key must still equal the oracle's sequential insertion"""
    rng = np.random.default_rng(11)
    n = 1_400_000
    xyz = (rng.random((n, 3)) * np.array([1.0, 1.0, 0.3])).astype(np.float32)             # a dense patch ...
    far = {5: (2.5, 0.2, 0.1), 300_000: (-3.0, 0.5, 0.2), 600_001: (0.5, 9.0, 0.1), 700_000: (0.2, -20.0, 0.3),
           980_000: (45.0, 1.0, 0.2), 1_250_000: (0.5, 0.5, 70.0), 1_399_999: (-150.0, 2.0, 1.0)}  # ... and late outliers
    for i, p in far.items():
        xyz[i] = p
    xyz[17] = (np.nan, 0.0, 0.0)
    g = gpu_stages(xyz)
    r = oracle.run(xyz, math=1)
    assert r.stats["growth_events"] >= 10
    _compare_vgs(xyz, g, r)


def test_origin_depth_limit_is_reported(built_lib):
    """an extent beyond 2^21 voxels per axis fails loudly (VGS_ERR_LIMIT), from the device loop as from the host one"""
    xyz = np.array([[0, 0, 0], [1, 1, 1], [4.0e5, 0, 0]], np.float32)
    h = capi_handle()
    try:
        h.set_points(xyz)
        with pytest.raises(Exception) as e:
            h.voxelize(0.15)
        assert "depth" in str(e.value)
    finally:
        h.close()
