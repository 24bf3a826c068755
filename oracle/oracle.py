"""ctypes wrapper around oracle/_build/libvgs_oracle.so (the CPU ORACLE).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED (see oracle/vgs_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvgs_oracle.so")


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("voxel_size", C.c_float),
        ("graph_size", C.c_float),
        ("sig_p", C.c_float), ("sig_n", C.c_float), ("sig_o", C.c_float),
        ("sig_e", C.c_float), ("sig_c", C.c_float), ("sig_w", C.c_float),
        ("cut_thred", C.c_float),
        ("points_min", C.c_int32), ("adjacency_min", C.c_int32), ("voxels_min", C.c_int32),
        ("leaf_order", C.c_int32), ("math", C.c_int32), ("near_tol", C.c_float),
    ]


# Task_File_VGS.txt:29-51 / Task_File_SVGS.txt:29-61 parameter sets
VGS_DEFAULT = dict(mode=0, voxel_size=0.15, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2,
                   sig_c=0.2, sig_w=2.0, cut_thred=0.3, points_min=10, adjacency_min=3, voxels_min=3,
                   leaf_order=0, math=1, near_tol=0.0)
SVGS_DEFAULT = dict(mode=1, voxel_size=0.05, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2,
                    sig_c=0.75, sig_w=1.0, cut_thred=0.5, points_min=10, adjacency_min=3, voxels_min=3,
                    leaf_order=0, math=1, near_tol=0.0)

KINDS = dict(
    BBOX=(0, np.float64), POINT_KEY=(1, np.uint32), POINT_UNIT=(2, np.int32), UNIT_KEY=(3, np.uint32),
    UNIT_CENTER=(4, np.float32), UNIT_OFFSETS=(5, np.int64), UNIT_POINTS=(6, np.int32),
    CENTROID=(7, np.float32), NORMAL=(8, np.float32), EIGEN=(9, np.float32), USED=(10, np.uint8),
    ADJ_OFFSETS=(11, np.int64), ADJ_IDX=(12, np.int32),
    CONN0_OFFSETS=(13, np.int64), CONN0_IDX=(14, np.int32),
    CONN1_OFFSETS=(15, np.int64), CONN1_IDX=(16, np.int32),
    CONN2_OFFSETS=(17, np.int64), CONN2_IDX=(18, np.int32),
    UNIT_CLUSTER=(19, np.int32), POINT_LABEL=(20, np.int32),
    CLUSTER_OFFSETS=(21, np.int64), CLUSTER_POINTS=(22, np.int32),
    NEAR_EDGES=(23, np.int32), STATS=(24, np.int64), ATTACH=(25, np.int32),
)
STAT_NAMES = ["n_points", "n_finite", "n_units", "n_used", "n_adjacency", "pair_evals", "nan_weights",
              "ub_corner", "near_threshold", "n_clusters_all", "n_clusters_exported", "octree_depth",
              "singles", "attached", "growth_events", "max_n_or_voxels"]

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no reference sources are used)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("vgs_oracle.cpp", "vccs_oracle.cpp", "vgs_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.vgso_create.restype = C.c_void_p
        L.vgso_create.argtypes = [C.POINTER(Params)]
        L.vgso_destroy.argtypes = [C.c_void_p]
        L.vgso_run.restype = C.c_int
        L.vgso_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int32]
        L.vgso_get.restype = C.c_void_p
        L.vgso_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.vgso_pair.argtypes = [C.POINTER(Params)] + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] * 2 + [C.c_void_p]
        L.vgso_eigen33.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vgso_features.argtypes = [C.POINTER(Params), C.c_void_p, C.c_int64, C.c_void_p]
        L.vgso_set_threads.argtypes = [C.c_int]
        L.vgso_set_threads.restype = C.c_int
        L.vgso_cut.restype = C.c_int
        L.vgso_cut.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.vgso_vccs.restype = C.c_int
        L.vgso_vccs.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(VccsParams), C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def set_threads(n: int):
    """threads of the per-unit local-graph loop (1 = as the reference runs; results do not depend on it)"""
    return int(lib().vgso_set_threads(int(n)))


def make_params(**kw) -> Params:
    base = dict(SVGS_DEFAULT if kw.get("mode", 0) == 1 else VGS_DEFAULT)
    base.update(kw)
    return Params(**base)


class Result(dict):
    """dict of numpy arrays (copies) keyed by the lower-case blob names + 'stats' dict."""
    __getattr__ = dict.__getitem__


def run(xyz: np.ndarray, labels: np.ndarray | None = None, max_label: int = 0, **kw) -> Result:
    """Run the full oracle pipeline.  xyz: (N,3) or (N,4) float32."""
    L = lib()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
    p = make_params(**kw)
    if labels is not None:
        labels = np.ascontiguousarray(labels, dtype=np.int32)
    h = L.vgso_create(C.byref(p))
    try:
        rc = L.vgso_run(h, xyz.ctypes.data, xyz.shape[0], xyz.shape[1],
                        labels.ctypes.data if labels is not None else None, max_label)
        if rc != 0:
            raise RuntimeError(f"oracle failed rc={rc}")
        out = Result()
        for name, (kind, dt) in KINDS.items():
            cnt = C.c_int64(0)
            ptr = L.vgso_get(h, kind, C.byref(cnt))
            if cnt.value == 0 or not ptr:
                arr = np.zeros(0, dtype=dt)
            else:
                arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))),
                                            shape=(cnt.value,)).copy()
            out[name.lower()] = arr
        for k3 in ("point_key", "unit_key", "unit_center", "centroid", "normal", "near_edges"):
            out[k3] = out[k3].reshape(-1, 3)
        out["eigen"] = out["eigen"].reshape(-1, 8)
        out["stats"] = dict(zip(STAT_NAMES, (int(v) for v in out["stats"])))
        return out
    finally:
        L.vgso_destroy(h)


def pair(c1, n1, e1, c2, n2, e2, flags1=7, flags2=7, **kw):
    """(S, A, T, E, C, w) of the ordered pair (v1, v2)."""
    L = lib()
    p = make_params(**kw)
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (c1, n1, e1, c2, n2, e2)]
    out = np.zeros(6, dtype=np.float32)
    L.vgso_pair(C.byref(p), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, flags1,
                a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data, flags2, out.ctypes.data)
    return out


def eigen33(mat, math=1):
    L = lib()
    m = np.ascontiguousarray(mat, dtype=np.float32).reshape(9)
    ev = np.zeros(3, np.float32)
    evec = np.zeros(9, np.float32)
    L.vgso_eigen33(m.ctypes.data, ev.ctypes.data, evec.ctypes.data, math)
    return ev, evec.reshape(3, 3)


def features(xyz, **kw):
    L = lib()
    p = make_params(**kw)
    x = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(14, np.float32)
    L.vgso_features(C.byref(p), x.ctypes.data, x.shape[0], out.ctypes.data)
    return out[:3], out[3:6], out[6:]


class VccsParams(C.Structure):
    _fields_ = [("voxel_res", C.c_float), ("seed_res", C.c_float), ("color_importance", C.c_float),
                ("spatial_importance", C.c_float), ("normal_importance", C.c_float), ("refine_iterations", C.c_int32),
                ("schedule", C.c_int32)]


def vccs(xyz, voxel_res=0.05, seed_res=0.25, color_importance=0.0, spatial_importance=0.25, normal_importance=0.75,
         refine_iterations=5, schedule=1, leaf_order=0) -> Result:
    """Supervoxel generator restatement (vccs_oracle.cpp) on the oracle's own voxel table.  Defaults are
    Task_File_SVGS.txt's.  Returns point_label, max_label, n_seeds, vox_normal (initial), vox_label and the voxel table."""
    L = lib()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    # voxel table only: no voxel is "used" and the graph radius reaches nobody, so the later stages are empty
    base = run(xyz, mode=0, voxel_size=voxel_res, graph_size=voxel_res * 0.5, points_min=2 ** 30, leaf_order=leaf_order, math=1)
    V = int(base.stats["n_units"])
    p = VccsParams(voxel_res, seed_res, color_importance, spatial_importance, normal_importance, refine_iterations, schedule)
    key = np.ascontiguousarray(base.unit_key, dtype=np.uint32)
    off = np.ascontiguousarray(base.unit_offsets, dtype=np.int64)
    pts = np.ascontiguousarray(base.unit_points, dtype=np.int32)
    origin = np.ascontiguousarray(base.bbox[:3], dtype=np.float64)
    lab = np.zeros(xyz.shape[0], np.int32)
    ml = C.c_int32(0)
    vn = np.zeros((V, 3), np.float32)
    vl = np.zeros(V, np.int32)
    ns = L.vgso_vccs(xyz.ctypes.data, xyz.shape[0], xyz.shape[1], V, key.ctypes.data, off.ctypes.data, pts.ctypes.data,
                     origin.ctypes.data, C.byref(p), lab.ctypes.data, C.byref(ml), vn.ctypes.data, vl.ctypes.data)
    return Result(point_label=lab, max_label=int(ml.value), n_seeds=int(ns), vox_normal=vn, vox_label=vl, unit_key=key,
                  unit_offsets=off, unit_points=pts, bbox=base.bbox, point_unit=base.point_unit)


def cut(w, cut_thred):
    L = lib()
    w = np.ascontiguousarray(w, dtype=np.float32)
    n = w.shape[0]
    out = np.zeros(n, np.int32)
    m = L.vgso_cut(cut_thred, w.ctypes.data, n, out.ctypes.data)
    return out[:m].copy()
