"""not-gpu: the drop-in I/O layer (row f2): PCD v0.7 ascii / binary / binary_compressed readers and
the positional task-file reader (IO.cpp:147-169) — pure host code, no GPU involved."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "pcd_roundtrip")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                    os.path.join(ROOT, "tests", "cpp", "pcd_roundtrip.cpp")], check=True)
    return EXE


def lzf_compress(data: bytes) -> bytes:
    """small greedy LZF encoder (literal runs + back references), enough to exercise the decoder"""
    out = bytearray()
    lit = bytearray()
    i, n = 0, len(data)
    last = {}

    def flush():
        for k in range(0, len(lit), 32):
            chunk = lit[k:k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit.clear()
    while i < n:
        key = data[i:i + 3]
        j = last.get(key, -1) if len(key) == 3 else -1
        last[key] = i
        if j >= 0 and 0 < i - j <= 8192:
            ln = 3
            while i + ln < n and ln < 264 and data[j + ln] == data[i + ln]:
                ln += 1
            flush()
            off = i - j - 1
            l2 = ln - 2
            if l2 < 7:
                out.append((l2 << 5) | (off >> 8))
            else:
                out.append((7 << 5) | (off >> 8))
                out.append(l2 - 7)
            out.append(off & 255)
            i += ln
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def _cloud(n=3000, seed=0):
    rng = np.random.default_rng(seed)
    xyz = rng.normal(size=(n, 3)).astype(np.float32) * 10
    xyz[::7] = xyz[0]          # repeats, so that the LZF stream contains back references
    return xyz


def _header(n, data, with_intensity=True):
    f = "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n" if with_intensity else \
        "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n" + f +
            f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {data}\n").encode()


def _read_back(exe, path, tmp_path):
    out = tmp_path / "out.f32"
    r = subprocess.run([exe, str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(out, np.float32).reshape(-1, 3)


def test_pcd_binary(exe, tmp_path):
    xyz = _cloud()
    rec = np.concatenate([xyz, np.full((len(xyz), 1), 3.5, np.float32)], 1)
    p = tmp_path / "b.pcd"
    p.write_bytes(_header(len(xyz), "binary") + rec.tobytes())
    np.testing.assert_array_equal(_read_back(exe, p, tmp_path), xyz)


def test_pcd_binary_compressed(exe, tmp_path):
    xyz = _cloud()
    inten = np.full(len(xyz), 3.5, np.float32)
    soa = xyz[:, 0].tobytes() + xyz[:, 1].tobytes() + xyz[:, 2].tobytes() + inten.tobytes()   # field-major
    comp = lzf_compress(soa)
    assert len(comp) < len(soa)          # back references were produced
    p = tmp_path / "c.pcd"
    p.write_bytes(_header(len(xyz), "binary_compressed") + np.array([len(comp), len(soa)], np.uint32).tobytes() + comp)
    np.testing.assert_array_equal(_read_back(exe, p, tmp_path), xyz)


def test_pcd_ascii(exe, tmp_path):
    xyz = _cloud(500)
    body = "\n".join(f"{x:.9g} {y:.9g} {z:.9g}" for x, y, z in xyz.tolist()) + "\n"
    p = tmp_path / "a.pcd"
    p.write_bytes(_header(len(xyz), "ascii", with_intensity=False) + body.encode())
    np.testing.assert_array_equal(_read_back(exe, p, tmp_path), xyz)


def test_pcd_ascii_nan_row_stays_nan(exe, tmp_path):
    """PCL writes `nan nan nan` for non-finite points in DATA ascii files; the reader must hand them on as NaN (the octree
    skips them, as addPointsFromInputCloud does) instead of inventing a point at the origin"""
    p = tmp_path / "n.pcd"
    p.write_bytes(_header(3, "ascii", with_intensity=False) + b"1 2 3\nnan nan nan\n4 5 6\n")
    got = _read_back(exe, p, tmp_path)
    assert got.shape == (3, 3)
    np.testing.assert_array_equal(got[[0, 2]], np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    assert np.isnan(got[1]).all()


def test_task_file_lines_are_positional(exe, tmp_path):
    """IO.cpp:162-165: every line (comments and blanks included) is kept; parameter k = line k; CRLF files"""
    lines = ["// header", "", "Seg", "//Tasks", "0.15", ""]
    p = tmp_path / "t.txt"
    p.write_bytes(("\r\n".join(lines) + "\r\n").encode())
    r = subprocess.run([exe, "task", str(p)], capture_output=True, text=True)
    got = r.stdout.splitlines()
    assert got[0] == str(len(lines)) and got[1:] == [f"[{l}]" for l in lines]


def _ply_header(fmt, n, props, nfaces=0, crlf=False):
    lines = ["ply", f"format {fmt} 1.0", "comment test", f"element vertex {n}"] + [f"property {t} {name}" for t, name in props]
    if nfaces:
        lines += [f"element face {nfaces}", "property list uchar int vertex_indices"]
    lines.append("end_header")
    return (("\r\n" if crlf else "\n").join(lines) + ("\r\n" if crlf else "\n")).encode()


def _read_ply(exe, path, tmp_path):
    out = tmp_path / "out.f32"
    r = subprocess.run([exe, "ply", str(path), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(out, np.float32).reshape(-1, 3)


def test_ply_ascii_with_colour_and_faces(exe, tmp_path):
    """inputPointCloudData2 (IO.h:83-97): x y z taken from among other scalar properties; faces ignored"""
    xyz = _cloud(400)
    body = "\n".join(f"{x:.9g} {y:.9g} {z:.9g} 10 20 30" for x, y, z in xyz.tolist()) + "\n3 0 1 2\n"
    p = tmp_path / "a.ply"
    p.write_bytes(_ply_header("ascii", len(xyz), [("float", "x"), ("float", "y"), ("float", "z"), ("uchar", "red"),
                                                   ("uchar", "green"), ("uchar", "blue")], nfaces=1) + body.encode())
    np.testing.assert_array_equal(_read_ply(exe, p, tmp_path), xyz)


@pytest.mark.parametrize("endian", ["little", "big"])
def test_ply_binary_mixed_properties(exe, tmp_path, endian):
    """binary PLY, both byte orders: an intensity before x, double z, a uchar after; CRLF header"""
    xyz = _cloud(1000)
    e = "<" if endian == "little" else ">"
    rec = np.zeros(len(xyz), dtype=[("intensity", e + "f4"), ("x", e + "f4"), ("y", e + "f4"), ("z", e + "f8"), ("cls", "u1")])
    rec["intensity"], rec["x"], rec["y"], rec["z"], rec["cls"] = 7.0, xyz[:, 0], xyz[:, 1], xyz[:, 2].astype(np.float64), 3
    p = tmp_path / "b.ply"
    p.write_bytes(_ply_header(f"binary_{endian}_endian", len(xyz), [("float", "intensity"), ("float", "x"), ("float", "y"),
                                                                   ("double", "z"), ("uchar", "cls")], crlf=True) + rec.tobytes())
    np.testing.assert_array_equal(_read_ply(exe, p, tmp_path), xyz)


def test_ply_errors(exe, tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"plx\n")
    assert subprocess.run([exe, "ply", str(p), str(tmp_path / "o")], capture_output=True, text=True).stderr.strip() == "rc -2"
    p.write_bytes(_ply_header("ascii", 2, [("float", "x"), ("float", "y")]) + b"1 2\n3 4\n")
    assert subprocess.run([exe, "ply", str(p), str(tmp_path / "o")], capture_output=True, text=True).stderr.strip() == "rc -3"
    p.write_bytes(_ply_header("binary_little_endian", 5, [("float", "x"), ("float", "y"), ("float", "z")]) + b"\0" * 20)
    assert subprocess.run([exe, "ply", str(p), str(tmp_path / "o")], capture_output=True, text=True).stderr.strip() == "rc -4"
    assert subprocess.run([exe, "ply", str(tmp_path / "missing.ply"), str(tmp_path / "o")], capture_output=True, text=True).stderr.strip() == "rc -1"


def test_output_point_cloud_data_roundtrip(exe, tmp_path):
    """outputPointCloudData (IO.h:100-108) writes a binary XYZ PCD that inputPointCloudData reads back"""
    xyz = _cloud(700)
    a = tmp_path / "a.pcd"
    a.write_bytes(_header(len(xyz), "binary", with_intensity=False) + xyz.tobytes())
    b = tmp_path / "b.pcd"
    r = subprocess.run([exe, "xyzpcd", str(a), str(b)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    np.testing.assert_array_equal(_read_back(exe, b, tmp_path), xyz)
