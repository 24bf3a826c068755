"""ctypes binding of libvgs_b200.so (include/vgs_b200.h).  Thin: every method is one C-ABI call.

The shared library is the product; this module only marshals numpy / raw device pointers into it.
There is no CPU fallback: if the library is missing, or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvgs_b200.so")

VGS_MODE_VGS, VGS_MODE_SVGS = 0, 1
STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "STATE", 4: "LIMIT", 5: "NO_DEVICE"}


class VgsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"vgs_b200 status {status} ({STATUS.get(status, '?')}): {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [("mode", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p), ("leaf_order", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class Sigmas(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("sig_p", "sig_n", "sig_o", "sig_e", "sig_c", "sig_w")]


class Params(C.Structure):
    _fields_ = [("voxel_size", C.c_float), ("graph_size", C.c_float), ("sig", Sigmas), ("cut_thred", C.c_float),
                ("points_min", C.c_int32), ("adjacency_min", C.c_int32), ("voxels_min", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("h2d_ms", "origin_ms", "voxelize_ms", "features_ms", "adjacency_ms", "graph_ms",
                                         "mutual_ms", "closest_ms", "components_ms", "labels_ms", "d2h_ms", "total_ms")] + \
               [("kernel_launches", C.c_int64), ("pair_cache_ms", C.c_float), ("reserved", C.c_float * 3)]


class KernelTiming(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_float), ("launches", C.c_int32), ("alg_bytes", C.c_int64), ("reserved", C.c_int64)]


class Counts(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_points", "n_finite", "n_voxels", "n_units", "n_used", "n_adjacency", "n_pairs",
                                         "n_singles", "n_attached", "n_clusters_all", "n_clusters_exported", "octree_depth",
                                         "closest_rounds", "max_neighbours")] + [("reserved", C.c_int64 * 2)]


class GroupCounts(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_ranks", "n_points", "n_tile_points", "n_tile_voxels", "n_adjacency", "n_singles",
                                         "n_clusters_exported", "n_cross_pairs", "octree_depth", "halo", "axis", "origin_rounds",
                                         "closest_rounds")] + [("cuts", C.c_int64 * 17), ("reserved", C.c_int64 * 2)]


class GroupTimings(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("origin_ms", "cuts_ms", "route_ms", "tiles_ms", "low_ms", "closest_ms", "components_ms",
                                         "merge_ms", "labels_ms", "total_ms")] + [("kernel_launches", C.c_int64), ("reserved", C.c_float * 4)]


BLOBS = dict(POINT_KEY=(1, np.uint32), POINT_UNIT=(2, np.int32), UNIT_KEY=(3, np.uint32), UNIT_CENTER=(4, np.float32),
             UNIT_OFFSETS=(5, np.int64), UNIT_POINTS=(6, np.int32), RECORDS=(7, np.float32), ADJ_OFFSETS=(11, np.int64),
             ADJ_IDX=(12, np.int32), CONN0_COUNT=(13, np.int32), CONN0_IDX=(14, np.int32), CONN1_COUNT=(15, np.int32),
             CONN1_IDX=(16, np.int32), ATTACH=(17, np.int32), UNIT_ROOT=(19, np.int32))

EXPORTED = ["vgs_create", "vgs_destroy", "vgs_acquire", "vgs_release", "vgs_pool_trim", "vgs_get_unit_adjacency", "vgs_last_error", "vgs_device_count", "vgs_set_points", "vgs_voxelize",
            "vgs_get_bounding_box", "vgs_set_bounding_box", "vgs_voxel_count", "vgs_get_voxel_centers", "vgs_set_supervoxel_labels", "vgs_make_supervoxels_grid", "vgs_make_supervoxels_vccs", "vgs_get_supervoxel_labels", "vgs_unit_count",
            "vgs_compute_features", "vgs_find_adjacency", "vgs_segment", "vgs_cluster_count", "vgs_get_point_labels",
            "vgs_get_clusters_csr", "vgs_run", "vgs_get_counts", "vgs_stage_timings", "vgs_debug_get", "vgs_kernel_timings",
            "vgs_group_unique_id", "vgs_group_create_nccl", "vgs_group_create_local", "vgs_group_destroy", "vgs_group_last_error",
            "vgs_group_run", "vgs_group_get_counts", "vgs_group_get_timings", "vgs_group_handle", "vgs_slab_choose_cuts"]

_lib = None


def load():
    """Load libvgs_b200.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a).  vgs_svgs_segmentation_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        L.vgs_last_error.restype = C.c_char_p
        L.vgs_last_error.argtypes = [C.c_void_p]
        L.vgs_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.vgs_destroy.argtypes = [C.c_void_p]
        L.vgs_acquire.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.vgs_release.argtypes = [C.c_void_p]
        L.vgs_release.restype = None
        L.vgs_pool_trim.argtypes = []
        L.vgs_pool_trim.restype = None
        L.vgs_get_unit_adjacency.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.vgs_device_count.argtypes = [C.POINTER(C.c_int)]
        L.vgs_set_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]
        L.vgs_voxelize.argtypes = [C.c_void_p, C.c_float]
        L.vgs_get_bounding_box.argtypes = [C.c_void_p, C.c_void_p]
        L.vgs_set_bounding_box.argtypes = [C.c_void_p, C.c_void_p]
        L.vgs_voxel_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.vgs_get_voxel_centers.argtypes = [C.c_void_p, C.c_void_p]
        L.vgs_set_supervoxel_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int]
        L.vgs_make_supervoxels_grid.argtypes = [C.c_void_p, C.c_float]
        L.vgs_make_supervoxels_vccs.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.vgs_get_supervoxel_labels.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.vgs_unit_count.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.vgs_compute_features.argtypes = [C.c_void_p, C.c_int]
        L.vgs_find_adjacency.argtypes = [C.c_void_p, C.c_float]
        L.vgs_segment.argtypes = [C.c_void_p, C.POINTER(Sigmas), C.c_float, C.c_int]
        L.vgs_cluster_count.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.vgs_get_point_labels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.vgs_get_clusters_csr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
        L.vgs_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int]
        L.vgs_get_counts.argtypes = [C.c_void_p, C.POINTER(Counts)]
        L.vgs_stage_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.vgs_debug_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
        L.vgs_kernel_timings.argtypes = [C.c_void_p, C.POINTER(KernelTiming), C.POINTER(C.c_int)]
        L.vgs_group_unique_id.argtypes = [C.c_void_p]
        L.vgs_group_create_nccl.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_int, C.c_int, C.c_void_p]
        L.vgs_group_create_local.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_int]
        L.vgs_group_destroy.argtypes = [C.c_void_p]
        L.vgs_group_last_error.restype = C.c_char_p
        L.vgs_group_last_error.argtypes = [C.c_void_p]
        L.vgs_group_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vgs_group_get_counts.argtypes = [C.c_void_p, C.POINTER(GroupCounts)]
        L.vgs_group_get_timings.argtypes = [C.c_void_p, C.POINTER(GroupTimings)]
        L.vgs_group_handle.restype = C.c_void_p
        L.vgs_group_handle.argtypes = [C.c_void_p, C.c_int]
        L.vgs_slab_choose_cuts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]
        _lib = L
    return _lib


def choose_cuts(hist: np.ndarray, shift: int, nranks: int):
    """host-only: (axis, cuts[nranks + 1]) the slab group derives from per-axis key histograms hist[3, nbins]"""
    L = load()
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    assert hist.ndim == 2 and hist.shape[0] == 3
    axis = C.c_int(0)
    cuts = np.zeros(nranks + 1, np.int32)
    st = L.vgs_slab_choose_cuts(hist.ctypes.data, hist.shape[1], shift, nranks, C.byref(axis), cuts.ctypes.data)
    if st != 0:
        raise VgsError(st, "vgs_slab_choose_cuts: bad argument")
    return int(axis.value), cuts


class Group:
    """One scene on several ranks (vgs_group): `nccl_id` given -> NCCL group with ONE local rank (this process);
    else a loopback group with `nranks` local ranks on one device."""

    def __init__(self, nranks, rank=0, nccl_id: bytes | None = None, device=0, stream=None, leaf_order=0):
        self.L = load()
        self.g = C.c_void_p()
        if stream == 0:
            stream = 1
        cfg = Config(VGS_MODE_VGS, device, stream, leaf_order)
        if nccl_id is not None:
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            st = self.L.vgs_group_create_nccl(C.byref(self.g), C.byref(cfg), nranks, rank, buf)
            self.nlocal = 1
        else:
            st = self.L.vgs_group_create_local(C.byref(self.g), C.byref(cfg), nranks)
            self.nlocal = nranks
        if st != 0:
            raise VgsError(st, self.L.vgs_group_last_error(None).decode())
        self.nranks = nranks

    @staticmethod
    def unique_id() -> bytes:
        L = load()
        buf = C.create_string_buffer(128)
        st = L.vgs_group_unique_id(buf)
        if st != 0:
            raise VgsError(st, L.vgs_group_last_error(None).decode())
        return buf.raw

    def close(self):
        if self.g:
            self.L.vgs_group_destroy(self.g)
            self.g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise VgsError(st, self.L.vgs_group_last_error(self.g).decode())

    def run_ptrs(self, params: Params, xyz_ptrs, ns, stride_bytes, on_device, label_ptrs):
        """raw pointers, one per local rank"""
        k = self.nlocal
        assert len(xyz_ptrs) == k and len(ns) == k and len(label_ptrs) == k
        xp = (C.c_void_p * k)(*[C.c_void_p(int(p)) for p in xyz_ptrs])
        lp = (C.c_void_p * k)(*[C.c_void_p(int(p)) for p in label_ptrs])
        nn = (C.c_int64 * k)(*[int(n) for n in ns])
        self._ck(self.L.vgs_group_run(self.g, C.byref(params), xp, nn, stride_bytes, 1 if on_device else 0, lp))

    def run(self, params: Params, slices):
        """host numpy slices (one per local rank, in rank order) -> list of int32 label arrays"""
        slices = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3) for x in slices]
        labels = [np.empty(max(x.shape[0], 1), np.int32) for x in slices]
        self.run_ptrs(params, [x.ctypes.data if x.shape[0] else 0 for x in slices], [x.shape[0] for x in slices], 12, False,
                      [l.ctypes.data for l in labels])
        return [l[:x.shape[0]] for l, x in zip(labels, slices)]

    def counts(self) -> dict:
        c = GroupCounts()
        self._ck(self.L.vgs_group_get_counts(self.g, C.byref(c)))
        d = {k: getattr(c, k) for k, _ in GroupCounts._fields_ if k not in ("reserved", "cuts")}
        d["cuts"] = [int(c.cuts[i]) for i in range(self.nranks + 1)]
        return d

    def timings(self) -> dict:
        t = GroupTimings()
        self._ck(self.L.vgs_group_get_timings(self.g, C.byref(t)))
        return {k: getattr(t, k) for k, _ in GroupTimings._fields_ if k != "reserved"}

    def handle(self, local_rank=0) -> "Handle":
        """the rank's own handle (owned by the group): kernel_timings(), counts(), timings() of its tile"""
        h = Handle.__new__(Handle)
        h.L = self.L
        h.h = C.c_void_p(self.L.vgs_group_handle(self.g, local_rank))
        h.n = 0
        h._keep = None
        h.close = lambda: None          # not ours to destroy
        return h


def make_params(voxel_size=0.15, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2, sig_c=0.2, sig_w=2.0,
                cut_thred=0.3, points_min=10, adjacency_min=3, voxels_min=3, **_ignored) -> Params:
    return Params(voxel_size, graph_size, Sigmas(sig_p, sig_n, sig_o, sig_e, sig_c, sig_w), cut_thred, points_min,
                  adjacency_min, voxels_min)


class Handle:
    """One vgs_handle (one device, one stream)."""

    def __init__(self, mode=VGS_MODE_VGS, device=0, stream=None, leaf_order=0, pooled=False):
        """pooled=True: vgs_acquire / vgs_release (the device working set is parked between handles, as the drop-in classes do)"""
        self.L = load()
        self.h = C.c_void_p()
        if stream == 0:
            stream = 1   # torch's default stream is CUDA's legacy default stream: pass cudaStreamLegacy, not NULL (= "own stream")
        cfg = Config(mode, device, stream, leaf_order)
        self.pooled = bool(pooled)
        st = (self.L.vgs_acquire if pooled else self.L.vgs_create)(C.byref(self.h), C.byref(cfg))
        if st != 0:
            raise VgsError(st, self.L.vgs_last_error(None).decode())
        self.n = 0
        self._keep = None

    def close(self):
        if self.h:
            (self.L.vgs_release if self.pooled else self.L.vgs_destroy)(self.h)
            self.h = C.c_void_p()

    def unit_adjacency(self, unit: int) -> np.ndarray:
        """getOneVoxelAdjacency (VS.h:268): ids within graph_size of the unit, nearest first, itself first"""
        n = C.c_int(0)
        ids = np.empty(256, np.int32)
        self._ck(self.L.vgs_get_unit_adjacency(self.h, unit, ids.ctypes.data, 256, C.byref(n)))
        return ids[:n.value].copy()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise VgsError(st, self.L.vgs_last_error(self.h).decode())

    # --- input ---
    def set_points(self, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
        self._keep = xyz
        self.n = xyz.shape[0]
        self._ck(self.L.vgs_set_points(self.h, xyz.ctypes.data, self.n, xyz.shape[1] * 4, 0))

    def set_points_device(self, ptr: int, n: int, stride_bytes: int = 12):
        self.n = n
        self._ck(self.L.vgs_set_points(self.h, C.c_void_p(ptr), n, stride_bytes, 1))

    def set_points_host_ptr(self, ptr: int, n: int, stride_bytes: int = 12):
        self.n = n
        self._ck(self.L.vgs_set_points(self.h, C.c_void_p(ptr), n, stride_bytes, 0))

    def set_supervoxel_labels(self, labels: np.ndarray, max_label: int = 0):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        self._ck(self.L.vgs_set_supervoxel_labels(self.h, labels.ctypes.data, max_label, 0))

    def make_supervoxels_grid(self, seed_size):
        self._ck(self.L.vgs_make_supervoxels_grid(self.h, seed_size))

    def make_supervoxels_vccs(self, seed_resolution=0.25, color_importance=0.0, spatial_importance=0.25, normal_importance=0.75,
                              refine_iterations=5):
        self._ck(self.L.vgs_make_supervoxels_vccs(self.h, seed_resolution, color_importance, spatial_importance, normal_importance,
                                                  refine_iterations))

    def supervoxel_labels(self):
        lab = np.empty(self.n, np.int32)
        ml = C.c_int32(0)
        self._ck(self.L.vgs_get_supervoxel_labels(self.h, lab.ctypes.data, C.byref(ml), 0))
        return lab, int(ml.value)

    # --- stages ---
    def voxelize(self, voxel_size):
        self._ck(self.L.vgs_voxelize(self.h, voxel_size))

    def bounding_box(self):
        out = np.zeros(6, np.float64)
        self._ck(self.L.vgs_get_bounding_box(self.h, out.ctypes.data))
        return out

    def set_bounding_box(self, box6):
        b = np.ascontiguousarray(box6, dtype=np.float64)
        self._ck(self.L.vgs_set_bounding_box(self.h, b.ctypes.data))

    def voxel_count(self):
        v = C.c_int64()
        self._ck(self.L.vgs_voxel_count(self.h, C.byref(v)))
        return v.value

    def unit_count(self):
        v = C.c_int64()
        self._ck(self.L.vgs_unit_count(self.h, C.byref(v)))
        return v.value

    def compute_features(self, points_min):
        self._ck(self.L.vgs_compute_features(self.h, points_min))

    def find_adjacency(self, graph_size):
        self._ck(self.L.vgs_find_adjacency(self.h, graph_size))

    def segment(self, sig: Sigmas, cut_thred, adjacency_min):
        self._ck(self.L.vgs_segment(self.h, C.byref(sig), cut_thred, adjacency_min))

    def cluster_count(self, voxels_min):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.L.vgs_cluster_count(self.h, voxels_min, C.byref(a), C.byref(b)))
        return a.value, b.value

    def point_labels(self, voxels_min) -> np.ndarray:
        out = np.empty(self.n, np.int32)
        self._ck(self.L.vgs_get_point_labels(self.h, voxels_min, out.ctypes.data, 0))
        return out

    def clusters_csr(self, voxels_min):
        nc, nt = C.c_int64(), C.c_int64()
        self._ck(self.L.vgs_get_clusters_csr(self.h, voxels_min, C.byref(nc), C.byref(nt), None, None))
        off = np.zeros(nc.value + 1, np.int64)
        idx = np.zeros(max(nt.value, 1), np.int32)
        self._ck(self.L.vgs_get_clusters_csr(self.h, voxels_min, C.byref(nc), C.byref(nt), off.ctypes.data, idx.ctypes.data))
        return off, idx[:nt.value]

    def run(self, params: Params, labels_out=None, on_device=False):
        """Whole pipeline; labels_out: numpy int32 array (host) or raw device pointer (on_device)."""
        if labels_out is None:
            labels_out = np.empty(self.n, np.int32)
        ptr = labels_out if isinstance(labels_out, int) else labels_out.ctypes.data
        self._ck(self.L.vgs_run(self.h, C.byref(params), C.c_void_p(ptr), 1 if on_device else 0))
        return labels_out

    # --- introspection ---
    def counts(self) -> dict:
        c = Counts()
        self._ck(self.L.vgs_get_counts(self.h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in Counts._fields_ if k != "reserved"}

    def timings(self) -> dict:
        t = Timings()
        self._ck(self.L.vgs_stage_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in Timings._fields_ if k != "reserved"}

    def kernel_timings(self) -> list:
        """[{name, ms, launches, alg_bytes}] per kernel group of the last run"""
        arr = (KernelTiming * 32)()
        n = C.c_int(32)
        self._ck(self.L.vgs_kernel_timings(self.h, arr, C.byref(n)))
        return [dict(name=arr[i].name.decode(), ms=float(arr[i].ms), launches=int(arr[i].launches), alg_bytes=int(arr[i].alg_bytes))
                for i in range(n.value)]

    def blob(self, name: str) -> np.ndarray:
        kind, dt = BLOBS[name]
        sz = C.c_size_t(0)
        self._ck(self.L.vgs_debug_get(self.h, kind, None, C.byref(sz)))
        out = np.empty(sz.value // np.dtype(dt).itemsize, dt)
        if sz.value:
            self._ck(self.L.vgs_debug_get(self.h, kind, out.ctypes.data, C.byref(sz)))
        return out
