"""Rows f1/f2 of SURVEY.md §8: the runnable driver (task-file reader, PCD in/out, segmentationVGS /
segmentationSVGS call sequences of the reference's `test` snippet) on the drop-in classes.
not-gpu: compiles, parses task files like IO.cpp:147-169; gpu: results equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "vgs_driver")

VGS_VALUES = {24: "2", 28: "0.15", 30: "0.5", 32: "0.2", 34: "0.2", 36: "0.2", 38: "0.2", 40: "0.2", 42: "2", 44: "0.3",
              46: "10", 48: "3", 50: "3"}
SVGS_VALUES = {24: "3", 28: "0.05", 30: "0.25", 32: "0.5", 34: "0.2", 36: "0.2", 38: "0.2", 40: "0.2", 42: "0.2", 44: "1",
               46: "0", 48: "0.25", 50: "0.75", 52: "0.5", 54: "10", 56: "10", 58: "3", 60: "3"}


def write_task_file(path, values, in_dir, in_name, out_name, nlines=66):
    """positional task file: parameter k is line k (0-based), CRLF line ends like the shipped files"""
    lines = [f"// line {k}" for k in range(nlines)]
    lines[12], lines[15], lines[18], lines[21] = in_dir, in_name, in_dir, out_name
    for k, v in values.items():
        lines[k] = v
    with open(path, "wb") as f:
        f.write(("\r\n".join(lines) + "\r\n").encode())


def write_pcd_binary(path, xyz, extra_field=True):
    n = xyz.shape[0]
    if extra_field:   # an intensity field the loader must skip (PointXYZ keeps x, y, z only)
        rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4")])
        rec["intensity"] = 7.0
        hdr = "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    else:
        rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
        hdr = "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    with open(path, "wb") as f:
        f.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n" + hdr +
                 f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n").encode())
        f.write(rec.tobytes())


def read_colored_pcd(path):
    raw = open(path, "rb").read()
    k = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    hdr = raw[:k].decode()
    n = int([l for l in hdr.splitlines() if l.startswith("POINTS")][0].split()[1])
    rec = np.frombuffer(raw[k:k + 16 * n], dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4")])
    return rec


def _compile():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib = os.path.join(ROOT, "vgs_svgs_segmentation_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                    os.path.join(ROOT, "tools", "vgs_driver.cpp"), "-L" + lib, "-lvgs_b200", "-Wl,-rpath," + lib], check=True)


def test_driver_compiles_and_reports_bad_input(built_lib, tmp_path):
    _compile()
    r = subprocess.run([EXE, str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 2 and "too short" in r.stderr
    write_task_file(tmp_path / "t.txt", VGS_VALUES, str(tmp_path) + "/", "none.pcd", "o.pcd")
    r = subprocess.run([EXE, str(tmp_path / "t.txt")], capture_output=True, text=True)
    assert r.returncode == 3 and "cannot read" in r.stderr


def _clusters_from_output(rec):
    """clusters in file order (one colour run per cluster)"""
    change = np.flatnonzero(np.diff(rec["rgb"].astype(np.int64)) != 0) + 1
    bounds = np.concatenate([[0], change, [len(rec)]])
    return [rec[a:b] for a, b in zip(bounds[:-1], bounds[1:])]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["vgs", "svgs"])
def test_driver_end_to_end(built_lib, tmp_path, mode):
    from oracle import oracle
    from vgs_svgs_segmentation_b200 import scenes
    _compile()
    xyz = scenes.construction_site(120_000, seed=6, extent=8.0)
    write_pcd_binary(tmp_path / "in.pcd", xyz)
    vals = VGS_VALUES if mode == "vgs" else SVGS_VALUES
    write_task_file(tmp_path / "task.txt", vals, str(tmp_path) + "/", "in.pcd", "out.pcd")
    cmd = [EXE, str(tmp_path / "task.txt")]
    if mode == "svgs":
        labels = scenes.supervoxel_labels_grid(xyz, 0.25)
        labels.tofile(tmp_path / "labels.i32")
        cmd += [str(tmp_path / "in.pcd"), str(tmp_path / "out.pcd"), str(tmp_path / "labels.i32")]
        ref = oracle.run(xyz, labels=labels, max_label=int(labels.max()) + 1, mode=1, math=1)
    else:
        ref = oracle.run(xyz, math=1)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    print(r.stdout)
    rec = read_colored_pcd(tmp_path / "out.pcd")
    got = _clusters_from_output(rec)
    off, pts = ref.cluster_offsets, ref.cluster_points
    assert len(got) == len(off) - 1 == ref.stats["n_clusters_exported"]
    for c, g in enumerate(got):          # same cluster order, same point sets
        exp = xyz[pts[off[c]:off[c + 1]]]
        assert len(g) == len(exp)
        a = np.stack([g["x"], g["y"], g["z"]], 1)
        assert np.array_equal(a[np.lexsort(a.T)], exp[np.lexsort(exp.T)])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["vgs", "svgs"])
def test_driver_display_exports(built_lib, tmp_path, mode):
    """VGS_DRIVER_EXPORT_DIR: the meshes / coloured clouds the snippet declares (test:42-47, 131-135) are written; SVGS runs
    without external labels, i.e. through the built-in supervoxel generator"""
    from oracle import oracle
    from vgs_svgs_segmentation_b200 import scenes
    from test_mesh_exports import read_ply
    _compile()
    xyz = scenes.two_planes(40_000)
    write_pcd_binary(tmp_path / "in.pcd", xyz, extra_field=False)
    write_task_file(tmp_path / "task.txt", VGS_VALUES if mode == "vgs" else SVGS_VALUES, str(tmp_path) + "/", "in.pcd", "out.pcd")
    exp = tmp_path / "exports"
    exp.mkdir()
    r = subprocess.run([EXE, str(tmp_path / "task.txt")], capture_output=True, text=True, env={**os.environ, "VGS_DRIVER_EXPORT_DIR": str(exp)})
    assert r.returncode == 0, r.stderr
    print(r.stdout)
    if mode == "vgs":
        ref = oracle.run(xyz, math=1)
        nu = int(ref.used.sum())
        for name, per_box in (("colored_voxels", 12), ("frames_voxels", 12), ("clustered_voxels", 12)):
            v, _, f = read_ply(exp / f"{name}.ply")
            assert len(v) == 8 * nu and len(f) == per_box * nu
        v, _, f = read_ply(exp / "normals_voxels.ply")
        assert len(v) == 2 * nu and len(f) == nu
        off = ref.unit_offsets
        assert len(read_colored_pcd(exp / "points_in_voxels.pcd")) == int((off[1:] - off[:-1])[ref.used > 0].sum())
    else:
        sv = oracle.vccs(xyz, schedule=1)
        rec = read_colored_pcd(exp / "points_in_supervoxels.pcd")
        # createSupervoxels drops label max_label (SV.h:313 loops k < max_label)
        assert len(rec) == int(((sv.point_label > 0) & (sv.point_label < sv.max_label)).sum())
        assert len(read_colored_pcd(exp / "points_in_voxels.pcd")) == len(xyz)
        v, _, f = read_ply(exp / "normals_supervoxels.ply")
        assert len(f) > 0 and len(v) == 2 * len(f)
