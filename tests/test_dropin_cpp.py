"""The C++ drop-in header (include/vgs_dropin/voxel_segmentation.h) driven exactly like the
reference's usage snippet (`test`:51-76).  not-gpu: it compiles and links against libvgs_b200.so and
fails loudly without a device; gpu: its labels equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "dropin_vgs")
P = ["0.15", "0.5", "0.2", "0.2", "0.2", "0.2", "0.2", "2", "0.3", "10", "3", "3"]   # Task_File_VGS.txt values


def _compile():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", EXE,
                    os.path.join(ROOT, "tests", "cpp", "dropin_vgs.cpp"), "-L" + os.path.join(ROOT, "vgs_svgs_segmentation_b200"),
                    "-lvgs_b200", "-Wl,-rpath," + os.path.join(ROOT, "vgs_svgs_segmentation_b200")], check=True)


def test_dropin_compiles_and_refuses_cpu(built_lib, tmp_path):
    import torch
    _compile()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    xyz = np.zeros((10, 3), np.float32)
    f = tmp_path / "x.f32"
    xyz.tofile(f)
    r = subprocess.run([EXE, str(f), "10", str(tmp_path / "l.i32")] + P, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_dropin_matches_oracle(built_lib, tmp_path):
    from oracle import oracle
    from vgs_svgs_segmentation_b200 import scenes
    _compile()
    xyz = scenes.construction_site(150_000, seed=4, extent=9.0)
    f = tmp_path / "x.f32"
    xyz.tofile(f)
    out = tmp_path / "l.i32"
    r = subprocess.run([EXE, str(f), str(xyz.shape[0]), str(out)] + P, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    print(r.stdout)
    lab = np.fromfile(out, np.int32)
    ref = oracle.run(xyz, math=1)
    np.testing.assert_array_equal(lab, ref.point_label)
    assert f"voxels {ref.stats['n_units']} " in r.stdout
    assert f"clusters_all {ref.stats['n_clusters_all']} exported {ref.stats['n_clusters_exported']} " in r.stdout


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_host_parallel_helpers(threads):
    """include/vgs_dropin/host_parallel.h: threaded list building / copying equals the serial definition (host only)"""
    exe = os.path.join(ROOT, "tests", "_build", "host_parallel_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "host_parallel_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, VGS_DROPIN_THREADS=threads))
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
