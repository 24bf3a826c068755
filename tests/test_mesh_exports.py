"""Display exports of the drop-in classes (reference voxel_segmentation.h:424-1104,
supervoxel_segmentation.h:424-611): cube / frame / normal meshes and coloured clouds, checked against
the oracle's voxel table, normals and clusters.  not-gpu: the program compiles; gpu: contents."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "mesh_exports")
FACES = [(0, 1, 2), (0, 2, 3), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (1, 2, 5), (2, 6, 5), (0, 3, 7), (0, 7, 4), (2, 3, 7), (2, 7, 6)]
EDGES = [(0, 1), (0, 3), (1, 2), (2, 3), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def _compile():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                    os.path.join(ROOT, "tests", "cpp", "mesh_exports.cpp"), "-L" + os.path.join(ROOT, "vgs_svgs_segmentation_b200"),
                    "-lvgs_b200", "-Wl,-rpath," + os.path.join(ROOT, "vgs_svgs_segmentation_b200")], check=True)


def read_ply(path):
    with open(path, "rb") as f:
        lines = f.read().decode().splitlines()
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0"
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    body = lines[lines.index("end_header") + 1:]
    assert len(body) == nv + nf
    v = np.array([l.split() for l in body[:nv]], dtype=np.float64).reshape(nv, 6)
    f = np.array([l.split() for l in body[nv:]], dtype=np.int64).reshape(nf, 4)
    assert (f[:, 0] == 3).all()
    return v[:, :3].astype(np.float32), v[:, 3:].astype(np.uint8), f[:, 1:]


def read_cloud(path):
    raw = np.fromfile(path, np.uint8).reshape(-1, 16)
    return raw[:, :12].copy().view(np.float32).reshape(-1, 3), raw[:, 12:15]


def pcl_centres(key, bbox, res):
    """OctreePointCloud::genLeafNodeCenterFromOctreeKey: double resolution, double box minimum."""
    return ((key.astype(np.float64) + 0.5) * float(np.float32(res)) + bbox[:3]).astype(np.float32)


def corners(c, res):
    h = 0.5 * float(np.float32(res))
    lo = (c.astype(np.float64) - h).astype(np.float32)
    hi = (c.astype(np.float64) + h).astype(np.float32)
    sel = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], bool)
    return np.where(sel[None], hi[:, None, :], lo[:, None, :])       # (n, 8, 3)


def test_mesh_exports_compile(built_lib):
    _compile()


@pytest.mark.gpu
def test_mesh_exports_match_oracle(built_lib, tmp_path):
    from oracle import oracle
    from vgs_svgs_segmentation_b200 import scenes
    _compile()
    xyz = scenes.construction_site(120_000, seed=8, extent=8.0)
    f = tmp_path / "x.f32"
    xyz.tofile(f)
    r = subprocess.run([EXE, str(f), str(xyz.shape[0]), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    print(r.stdout)
    ref = oracle.run(xyz, math=1)
    used = np.flatnonzero(ref.used)
    nu = len(used)
    ctr = pcl_centres(ref.unit_key[used], ref.bbox, 0.15)

    # drawColorMapofVoxels: 8 corners + 12 triangles per used voxel, voxel-id order, reference numbering
    v, c, faces = read_ply(tmp_path / "vgs_boxes.ply")
    assert len(v) == 8 * nu and len(faces) == 12 * nu
    np.testing.assert_array_equal(v.reshape(nu, 8, 3), corners(ctr, 0.15))
    np.testing.assert_array_equal(faces.reshape(nu, 12, 3), np.arange(nu)[:, None, None] * 8 + np.array(FACES)[None])
    assert (c.reshape(nu, 8, 3) == c.reshape(nu, 8, 3)[:, :1]).all()
    # drawFrameMapofVoxels: same vertices, 12 degenerate a-b-a triangles per box
    vf, _, ff = read_ply(tmp_path / "vgs_frames.ply")
    np.testing.assert_array_equal(vf, v)
    exp = np.array([(a, b, a) for a, b in EDGES])
    np.testing.assert_array_equal(ff.reshape(nu, 12, 3), np.arange(nu)[:, None, None] * 8 + exp[None])
    # drawNormofVoxels: centre -> centre + resolution * normal
    vn, _, fn = read_ply(tmp_path / "vgs_normals.ply")
    assert len(vn) == 2 * nu and len(fn) == nu
    np.testing.assert_array_equal(vn[0::2], ctr)
    res = np.float32(0.15)
    np.testing.assert_array_equal(vn[1::2], ctr + res * ref.normal[used])
    np.testing.assert_array_equal(fn, np.stack([np.arange(nu) * 2, np.arange(nu) * 2 + 1, np.arange(nu) * 2], 1))
    # drawColorMapofPointsinVoxels: the points of the used voxels, voxel by voxel
    pv, pc = read_cloud(tmp_path / "vgs_points_in_voxels.bin")
    off = ref.unit_offsets
    idx = np.concatenate([ref.unit_points[off[u]:off[u + 1]] for u in used])
    np.testing.assert_array_equal(pv, xyz[idx])
    sizes = (off[1:] - off[:-1])[used]
    starts = np.cumsum(sizes) - sizes
    assert (pc == np.repeat(pc[starts], sizes, axis=0)).all()
    # drawColorMapofClusteredVoxels: used voxels grouped by cluster (cluster order), one colour per cluster,
    # centres = voxel_centers_ (float-narrowed origin, VS.h:690)
    vc, cc, fc = read_ply(tmp_path / "vgs_clustered.ply")
    assert len(vc) == 8 * nu and len(fc) == 12 * nu
    cl = ref.unit_cluster[used]
    order = np.lexsort((used, cl))
    np.testing.assert_array_equal(vc.reshape(nu, 8, 3), corners(ref.unit_center[used[order]], 0.15))
    box_col = cc.reshape(nu, 8, 3)[:, 0]
    same_cluster = cl[order][1:] == cl[order][:-1]
    assert (box_col[1:][same_cluster] == box_col[:-1][same_cluster]).all()

    # SVGS with the built-in seed grid: replay the oracle on the generated supervoxels
    unit = np.fromfile(tmp_path / "svgs_point_unit.i32", np.int32)
    lab = np.fromfile(tmp_path / "svgs_labels.i32", np.int32)
    sref = oracle.run(xyz, labels=unit + 1, max_label=int(unit.max()) + 2, **{**oracle.SVGS_DEFAULT, "math": 1})
    np.testing.assert_array_equal(lab, sref.point_label)
    soff = sref.unit_offsets
    big = np.flatnonzero((soff[1:] - soff[:-1]) > 10)
    vs, _, fs = read_ply(tmp_path / "svgs_normals.ply")
    assert len(vs) == 2 * len(big) and len(fs) == len(big)
    np.testing.assert_array_equal(vs[0::2], sref.centroid[big])
    np.testing.assert_array_equal(vs[1::2], sref.centroid[big] + np.float32(0.25) * sref.normal[big])
    ps, pcs = read_cloud(tmp_path / "svgs_points_in_supervoxels.bin")
    np.testing.assert_array_equal(ps, xyz[sref.unit_points])
    pvx, pcv = read_cloud(tmp_path / "svgs_points_in_voxels.bin")
    assert len(pvx) == len(xyz)
    # voxel by voxel in descending x-major Morton order of the 0.05 m octree keys, points ascending inside a voxel
    vref = oracle.run(xyz, voxel_size=0.05, math=1)
    np.testing.assert_array_equal(pvx, xyz[vref.unit_points])
    vsz = vref.unit_offsets[1:] - vref.unit_offsets[:-1]
    vst = np.cumsum(vsz) - vsz
    assert (pcv == np.repeat(pcv[vst], vsz, axis=0)).all()
