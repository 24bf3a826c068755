"""not-gpu: the per-thread device functions of csrc/vgs_math.cuh, compiled for the host by
tests/hostcheck/hostcheck.cpp, must reproduce the oracle bit for bit (this is the arithmetic the
CUDA kernels execute; IEEE basic operations are identical on host and device)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle


@pytest.fixture(scope="module")
def hc(built_lib):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = C.CDLL(os.path.join(root, "tests", "_build", "libvgs_hostcheck.so"))
    L.hc_morton.restype = C.c_uint64
    return L


def _rec(c, n, e, fl):
    r = np.zeros(16, np.float32)
    if fl & 1: r[:3] = c
    if fl & 2: r[3:6] = n
    if fl & 4: r[6:14] = e
    r[15:16].view(np.int32)[0] = fl
    return r


def test_unit_record_matches_oracle(hc):
    rng = np.random.default_rng(0)
    for t in range(1500):
        cnt = int(rng.integers(1, 60))
        mode = t % 2
        kind = t % 5
        pts = rng.normal(size=(cnt, 3)).astype(np.float32) * np.array([0.05, 0.05, 0.004 if kind else 0.05], np.float32) \
            + rng.uniform(-30, 30, 3).astype(np.float32)
        if kind == 3:
            pts[:, 0] = pts[0, 0]      # degenerate (planar in x): quadratic-root path of eigen33
        rec = np.zeros(16, np.float32)
        hc.hc_unit_record(pts.ctypes.data, cnt, 1, mode, rec.ctypes.data)
        c, n, e = oracle.features(pts, mode=mode, math=1)
        np.testing.assert_array_equal(rec[:3].view(np.uint32), c.view(np.uint32))
        np.testing.assert_array_equal(rec[3:6].view(np.uint32), n.view(np.uint32))
        np.testing.assert_array_equal(rec[6:14].view(np.uint32), e.view(np.uint32))


def test_pair_weights_match_oracle_both_orders(hc):
    rng = np.random.default_rng(1)
    for t in range(6000):
        mode = t % 2

        def mk():
            c = rng.uniform(-30, 30, 3).astype(np.float32)
            n = rng.normal(size=3); n = (n / np.linalg.norm(n)).astype(np.float32)
            e = rng.uniform(0, 1, 8).astype(np.float32)
            fl = {0: 0, 1: 6, 2: 5, 3: 3}.get(int(rng.integers(0, 14)), 7)
            return c, n, e, fl
        c1, n1, e1, f1 = mk(); c2, n2, e2, f2 = mk()
        if t % 3 == 0:
            c2 = (c1 + rng.normal(size=3) * 0.2).astype(np.float32)
            n2 = n1 + rng.normal(size=3) * 0.05; n2 = (n2 / np.linalg.norm(n2)).astype(np.float32)
            e2 = (e1 + rng.normal(size=8) * 0.01).astype(np.float32)
        if t % 97 == 0:
            n2 = n1.copy()          # parallel normals: acos argument may exceed 1 -> NaN weight
        sg = np.array([0.2, 0.2, 0.2, 0.2, 0.2, 1.0 if mode else 2.0], np.float32)
        out = np.zeros(2, np.float32)
        ra, rb = _rec(c1, n1, e1, f1), _rec(c2, n2, e2, f2)   # keep the buffers alive across the call
        hc.hc_pair(ra.ctypes.data, rb.ctypes.data, sg.ctypes.data, mode, out.ctypes.data)
        kw = dict(mode=mode, math=1, sig_w=float(sg[5]), sig_c=0.2)
        exp = np.array([oracle.pair(c1, n1, e1, c2, n2, e2, f1, f2, **kw)[5], oracle.pair(c2, n2, e2, c1, n1, e1, f2, f1, **kw)[5]], np.float32)
        np.testing.assert_array_equal(out.view(np.uint32), exp.view(np.uint32))


def test_morton_roundtrip(hc):
    rng = np.random.default_rng(2)
    for _ in range(500):
        k = [int(x) for x in rng.integers(0, 1 << 21, 3)]
        m = hc.hc_morton(*k)
        ref = 0
        for b in range(20, -1, -1):
            ref = (ref << 3) | (((k[0] >> b) & 1) << 2) | (((k[1] >> b) & 1) << 1) | ((k[2] >> b) & 1)
        assert m == ref
        out = (C.c_uint32 * 3)()
        hc.hc_demorton(C.c_uint64(m), out)
        assert list(out) == k


def test_fast_acos_exp_equal_the_library_rounding():
    """csrc/vgs_math.cuh: the short float-rounded acos / exp evaluations (IEEE operations only: same bits on host and device)
    against (float)acos((double)x) / (float)exp(y) — every 64th float of [-1, 1], every 64th weight argument, 2e7 random
    exponents (tools/fastmath_check.cpp without an argument checks EVERY float: 6.5e9 evaluations, 0 differences)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "fastmath_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fopenmp", os.path.join(root, "tools", "fastmath_check.cpp"), "-o", exe]
    if "fma" in open("/proc/cpuinfo").read():
        cmd.insert(1, "-mfma")          # hardware fma: the same results as libm's exact software fma, much faster
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe, "quick"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
