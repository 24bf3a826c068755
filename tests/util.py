"""Shared helpers for the parity tests: run the CUDA path through the C ABI and collect every
stage's output in the oracle's layout."""
import numpy as np

from vgs_svgs_segmentation_b200 import capi

VGS_PARAMS = dict(voxel_size=0.15, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2, sig_c=0.2, sig_w=2.0,
                  cut_thred=0.3, points_min=10, adjacency_min=3, voxels_min=3)
SVGS_PARAMS = dict(voxel_size=0.05, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2, sig_c=0.75, sig_w=1.0,
                   cut_thred=0.5, points_min=10, adjacency_min=3, voxels_min=3)


def csr_sets(offsets, counts, idx):
    """list of sorted tuples from lists stored at `offsets` with `counts` valid entries each"""
    return [tuple(sorted(idx[o:o + c].tolist())) for o, c in zip(offsets[:-1].tolist(), counts.tolist())]


def gpu_stages(xyz, mode=0, labels=None, max_label=0, leaf_order=0, **kw):
    p = dict(SVGS_PARAMS if mode == 1 else VGS_PARAMS)
    p.update(kw)
    h = capi.Handle(mode=mode, leaf_order=leaf_order)
    try:
        h.set_points(xyz)
        h.voxelize(p["voxel_size"])
        out = dict(bbox=h.bounding_box(), n_voxels=h.voxel_count())
        if mode == 1:
            h.set_supervoxel_labels(labels, max_label)
        else:
            out["point_key"] = h.blob("POINT_KEY").reshape(-1, 3)
        h.compute_features(p["points_min"])
        out["point_unit"] = h.blob("POINT_UNIT")
        out["unit_offsets"] = h.blob("UNIT_OFFSETS")
        out["unit_points"] = h.blob("UNIT_POINTS")
        rec = h.blob("RECORDS").reshape(-1, 16)
        out["records"] = rec
        out["centroid"], out["normal"], out["eigen"] = rec[:, 0:3], rec[:, 3:6], rec[:, 6:14]
        out["count"] = rec[:, 14].view(np.int32)
        out["flags"] = rec[:, 15].view(np.int32)
        out["used"] = ((out["flags"] & 8) != 0).astype(np.uint8)
        if mode == 0:
            out["unit_key"] = h.blob("UNIT_KEY").reshape(-1, 3)
            out["unit_center"] = h.blob("UNIT_CENTER").reshape(-1, 3)
        h.find_adjacency(p["graph_size"])
        out["adj_offsets"] = h.blob("ADJ_OFFSETS")
        out["adj_idx"] = h.blob("ADJ_IDX")
        sig = capi.Sigmas(p["sig_p"], p["sig_n"], p["sig_o"], p["sig_e"], p["sig_c"], p["sig_w"])
        h.segment(sig, p["cut_thred"], p["adjacency_min"])
        out["conn0_count"] = h.blob("CONN0_COUNT")
        out["conn0_idx"] = h.blob("CONN0_IDX")
        out["conn1_count"] = h.blob("CONN1_COUNT")
        out["conn1_idx"] = h.blob("CONN1_IDX")
        out["attach"] = h.blob("ATTACH")
        out["unit_root"] = h.blob("UNIT_ROOT")
        out["n_clusters"] = h.cluster_count(p["voxels_min"])
        out["point_label"] = h.point_labels(p["voxels_min"])
        out["clusters_csr"] = h.clusters_csr(p["voxels_min"])
        out["counts"] = h.counts()
        out["timings"] = h.timings()
        return out
    finally:
        h.close()


def oracle_conn_sets(off, idx):
    return [tuple(idx[a:b].tolist()) for a, b in zip(off[:-1].tolist(), off[1:].tolist())]
