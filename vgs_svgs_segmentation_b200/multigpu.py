"""Multi-GPU orchestration (one process per GPU, torch.distributed for the plumbing).

Exact partition of the dominant stage (SURVEY.md §8e / DESIGN.md §6): every rank holds the same cloud
and runs the cheap stages (voxelise, features, adjacency) itself, so voxel ids, keys and adjacency are
identical everywhere; the local-graph stage (pair cache + per-voxel cut, ~85 % of a step) is split over
contiguous voxel-id ranges balanced by sum(n^2); ranks then exchange their connect-list slices (one
broadcast per rank over NCCL/NVLink) and each finishes the mutual filter, closest check and components on
the complete lists.  Labels are bit-identical to the single-GPU run by construction.
"""
from __future__ import annotations

import numpy as np


def unit_ranges(adj_offsets: np.ndarray, world: int):
    """contiguous unit-id ranges with ~equal sum of n^2 (n = neighbourhood size, the pair-work proxy)"""
    n = np.diff(adj_offsets).astype(np.float64)
    w = np.concatenate([[0.0], np.cumsum(n * n)])
    nu = len(n)
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(w, w[-1] * r / world)))
    cuts.append(nu)
    cuts = [min(max(c, 0), nu) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange_connect(ranges, slot_ranges, rank, export_fn, import_fn, new_tensor, broadcast):
    """Every rank ends up with all ranks' connect-list slices.
    ranges[r] = (first_unit, last_unit); slot_ranges[r] = (e_first, e_last) adjacency slots of that range;
    export_fn(first, last, cnt, idx) fills tensors from the local handle; import_fn stores them;
    new_tensor(n) allocates an int32 tensor on the exchange device; broadcast(t, src) is the collective."""
    for r, ((a, b), (e0, e1)) in enumerate(zip(ranges, slot_ranges)):
        cnt = new_tensor(b - a)
        idx = new_tensor(e1 - e0)
        if r == rank:
            export_fn(a, b, cnt, idx)
        broadcast(cnt, r)
        broadcast(idx, r)
        if r != rank:
            import_fn(a, b, cnt, idx)


def segment_partitioned(h, params, rank: int, world: int, labels_out=None, on_device=False):
    """Full pipeline on handle `h` (points already set) with stage 4+5a partitioned over `world` ranks.
    Needs an initialised torch.distributed process group (NCCL) when world > 1."""
    import torch
    import torch.distributed as dist
    from . import capi

    p = params
    h.voxelize(p.voxel_size)
    h.compute_features(p.points_min)
    h.find_adjacency(p.graph_size)
    sig = p.sig
    if world == 1:
        h.segment(sig, p.cut_thred, p.adjacency_min)
    else:
        ranges, slots = h.unit_ranges(world)      # same rule as unit_ranges() below, evaluated by the library
        a, b = ranges[rank]
        h.segment_partial(sig, p.cut_thred, a, b)
        dev = torch.device("cuda", torch.cuda.current_device())
        pool = h.__dict__.setdefault("_exchange_pool", {})   # exchange buffers are reused across steps

        counter = [0]

        def new_tensor(n):
            k = counter[0]
            counter[0] += 1
            t = pool.get(k)
            if t is None or t.numel() < n or t.device != dev:
                t = torch.empty(max(int(n) + int(n) // 8, 1), dtype=torch.int32, device=dev)
                pool[k] = t
            return t[:max(int(n), 1)]
        exchange_connect(
            ranges, slots, rank,
            export_fn=lambda f, l, c, i: h.export_connect(f, l, c.data_ptr(), i.data_ptr()),
            import_fn=lambda f, l, c, i: h.import_connect(f, l, c.data_ptr(), i.data_ptr()),
            new_tensor=new_tensor,
            broadcast=lambda t, src: dist.broadcast(t, src=src))
        h.segment_finish(sig, p.cut_thred, p.adjacency_min)
    if labels_out is None:
        return h.point_labels(p.voxels_min)
    ptr = labels_out if isinstance(labels_out, int) else labels_out.ctypes.data
    h._ck(h.L.vgs_get_point_labels(h.h, p.voxels_min, capi.C.c_void_p(ptr), 1 if on_device else 0))
    return labels_out
