"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, stage by stage.

Bar (BASELINE.json north_star): voxel keys, adjacency lists and canonical labels bit-exact;
per-unit features within 1e-5 relative (they are in fact compared bit for bit first and the number
of non-identical floats is reported)."""
import numpy as np
import pytest

from oracle import oracle
from vgs_svgs_segmentation_b200 import scenes

from util import VGS_PARAMS, csr_sets, gpu_stages, oracle_conn_sets

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
    """VGS_B200_NO_PAIR_CACHE=1: weights evaluated inside every local graph (the SVGS code path) must
    give the same lists and labels as the offset-indexed pair cache."""
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


def test_vgs_pair_cache_without_bitmap(built_lib, monkeypatch):
    """VGS_B200_NO_BITMAP=1: the pair cache finds partners by hash probes of every stencil offset (the path taken for
    octrees deeper than 11 levels) instead of reading z-runs of the occupancy bitmap; same lists and labels."""
    monkeypatch.setenv("VGS_B200_NO_BITMAP", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_NO_BITMAP")
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_single_stream_classes(built_lib, monkeypatch):
    """VGS_B200_CLASS_STREAMS=1: all size-class launches on the handle's stream (default: 6 streams, forked / joined with
    events); the connect lists do not depend on how the launches overlap."""
    monkeypatch.setenv("VGS_B200_CLASS_STREAMS", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_CLASS_STREAMS")
    _compare_vgs(xyz, g, oracle.run(xyz, math=1))


def test_vgs_adjacency_two_pass(built_lib, monkeypatch):
    """VGS_B200_ADJ_TWO_PASS=1: count / scan / probe-again adjacency (used when the staging rows of the one-pass
    variant would not fit) gives the same lists as the default one-pass build."""
    monkeypatch.setenv("VGS_B200_ADJ_TWO_PASS", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    monkeypatch.delenv("VGS_B200_ADJ_TWO_PASS")
    g2 = gpu_stages(xyz)
    assert np.array_equal(g["adj_offsets"], g2["adj_offsets"]) and np.array_equal(g["adj_idx"], g2["adj_idx"])
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


def test_full_size_properties_10m(built_lib):
    """BASELINE.json configs[2] at full size (10 M points): the oracle would need minutes, so the run is
    checked through size-independent properties: canonical labels (label = smallest point index of the
    cluster and that point carries it), every exported cluster has more than voxels_min voxels, cluster
    counts are consistent, a second run is bit-identical, and unlabelled points belong to dropped clusters."""
    from vgs_svgs_segmentation_b200 import capi
    xyz = scenes.construction_site(10_000_000, seed=1, extent=70.0)
    h = capi.Handle()
    h.set_points(xyz)
    p = capi.make_params(**VGS_PARAMS)
    lab = h.run(p).copy()
    c = h.counts()
    root = h.blob("UNIT_ROOT")
    pu = h.blob("POINT_UNIT")
    n_all, n_exp = h.cluster_count(VGS_PARAMS["voxels_min"])
    assert c["n_points"] == 10_000_000 and c["n_voxels"] == len(root) and n_all == len(np.unique(root))
    labelled = lab >= 0
    ids = np.unique(lab[labelled])
    assert len(ids) == n_exp
    assert np.array_equal(lab[ids], ids)                       # the smallest point of a cluster carries its own index
    first = np.full(lab.max() + 1, -1, np.int64)
    idx = np.flatnonzero(labelled)
    first[lab[idx][::-1]] = idx[::-1]                          # first occurrence of each label
    assert np.array_equal(first[ids], ids)                     # ... and no smaller point has that label
    # label is a function of the voxel's root, and exported roots have > voxels_min voxels
    r_of_point = root[pu]
    sizes = np.bincount(root, minlength=len(root))
    assert np.all(sizes[r_of_point[labelled]] > VGS_PARAMS["voxels_min"])
    assert np.all(sizes[r_of_point[~labelled]] <= VGS_PARAMS["voxels_min"])
    lab_of_root = np.full(len(root), -2, np.int64)
    lab_of_root[r_of_point] = lab
    assert np.array_equal(lab_of_root[r_of_point], lab)
    h.set_points(xyz)
    lab2 = h.run(p)
    h.close()
    assert np.array_equal(lab, lab2)


def test_vgs_cta_kernel_path(built_lib, monkeypatch):
    """VGS_B200_NO_WARP_KERNEL=1: the cached CTA-per-unit kernel (also the fallback of the warp-per-unit
    kernel) must give the oracle's lists and labels too."""
    monkeypatch.setenv("VGS_B200_NO_WARP_KERNEL", "1")
    xyz = _scene("site")
    g = gpu_stages(xyz)
    r = oracle.run(xyz, math=1)
    _compare_vgs(xyz, g, r)


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


def test_svgs_tiny(built_lib):
    rng = np.random.default_rng(2)
    xyz = (rng.random((500, 3)) * np.array([2.0, 2.0, 0.02])).astype(np.float32) + 0.5
    labels = scenes.supervoxel_labels_grid(xyz, 0.25)
    ml = int(labels.max()) + 1
    g = gpu_stages(xyz, mode=1, labels=labels, max_label=ml)
    r = oracle.run(xyz, labels=labels, max_label=ml, mode=1, math=1)
    _compare_svgs(xyz, labels, g, r)
