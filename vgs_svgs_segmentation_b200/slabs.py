"""ONE scene on N GPUs: plumbing around the C-ABI slab group (include/vgs_b200.h, vgs_group_*).

The split itself — global PCL origin, slab cuts, halo routing, per-tile pipeline, closest-check rounds, cross-slab
union-find, labels sent home — lives in libvgs_b200.so (csrc/vgs_group.inl, NCCL bound with dlopen).  This module
only (1) ships rank 0's NCCL unique id to the other ranks through torch.distributed (any backend: the CPU tests use
gloo), (2) says which points a rank holds, (3) is bench.py's N > 1 arm.
"""
from __future__ import annotations

import time

import numpy as np

from . import capi


def slice_bounds(n_total: int, world: int, rank: int):
    """[first, last) of the cloud held by `rank`: contiguous index ranges in rank order (the split of the reference's
    single input cloud, VS.h:94-102; point indices are global)"""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def share_unique_id(dist, rank: int, device=None) -> bytes:
    """rank 0 makes the 128-byte NCCL id (vgs_group_unique_id), every rank gets it by a broadcast on `dist`'s default group"""
    import torch
    raw = capi.Group.unique_id() if rank == 0 else bytes(128)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def create_group(dist, rank: int, world: int, device: int, stream=None) -> capi.Group:
    """collective: an NCCL slab group with this process as rank `rank`"""
    import torch
    backend = dist.get_backend()
    nccl_id = share_unique_id(dist, rank, torch.device("cuda", device) if backend == "nccl" else None)
    return capi.Group(world, rank=rank, nccl_id=nccl_id, device=device, stream=stream)


def scene_parts(cfg_name, configs, n_gpus):
    """(number of parts, points per part, extent): the workload of bench.py's slab arm.
    site10m / town2m: N parts of the config's size on sqrt(N) times the extent (weak scaling: per-GPU work fixed; N = 1 is
    exactly the single-GPU workload).  urban100m / city1b: always 8 / 80 parts of 12.5 M points on a 320 m / 1012 m ground
    (strong scaling: the same scene for every N)."""
    c = configs[cfg_name]
    if c["scene"] == "urban":
        parts = max(1, c["points"] // 12_500_000)
        return parts, c["points"] // parts, 320.0 * (c["points"] / 100_000_000) ** 0.5
    base = {"construction_site": 70.0, "town": 60.0}[c["scene"]]
    return n_gpus, c["points"], base * n_gpus ** 0.5


def make_part(cfg_name, configs, part, points, extent):
    from . import scenes
    c = configs[cfg_name]
    if c["scene"] == "construction_site":
        return scenes.construction_site(points, seed=1 + part, extent=extent)
    if c["scene"] == "town":
        return scenes.town(points, seed=20170610 + part, extent=extent)
    return scenes.urban(points, seed=2 + part, extent=extent)


def bench(args, cfg, pd, rank, world, local, flush, barrier, ClockSampler, measured_peak, CONFIGS):
    """bench.py --gpus N (N > 1), --mode slabs: ONE scene, one slab per rank.  Returns the JSON dict on rank 0."""
    import torch
    import torch.distributed as dist

    nparts, ppp, extent = scene_parts(cfg, CONFIGS, world)
    if args.points:
        ppp = max(1, args.points // nparts)
        if CONFIGS[cfg]["scene"] == "urban":
            extent = 320.0 * (args.points / 100_000_000) ** 0.5
    assert nparts % world == 0, "the scene's parts must divide over the ranks"
    my_parts = list(range(rank * nparts // world, (rank + 1) * nparts // world))
    pts = np.concatenate([make_part(cfg, CONFIGS, j, ppp, extent) for j in my_parts], axis=0)
    n = pts.shape[0]
    n_total = ppp * nparts
    stream = torch.cuda.current_stream()
    host_xyz = torch.from_numpy(pts).pin_memory()
    host_lab = torch.empty(n, dtype=torch.int32).pin_memory()
    dev_xyz = host_xyz.cuda()
    dev_lab = torch.empty(n, dtype=torch.int32, device="cuda")
    g = create_group(dist, rank, world, local, stream.cuda_stream)
    params = capi.make_params(**pd)

    def step_resident():
        g.run_ptrs(params, [dev_xyz.data_ptr()], [n], 12, True, [dev_lab.data_ptr()])

    def step_e2e():
        g.run_ptrs(params, [host_xyz.data_ptr()], [n], 12, False, [host_lab.data_ptr()])
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    h0 = g.handle(0)
    launches0 = g.timings()["kernel_launches"]
    times, gt_acc, kern_acc = [], {}, {}
    barrier()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_resident()
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
            for k, v in g.timings().items():
                gt_acc[k] = gt_acc.get(k, 0.0) + v
            for kt in h0.kernel_timings():
                a = kern_acc.setdefault(kt["name"], dict(ms=0.0, launches=kt["launches"], alg_bytes=kt["alg_bytes"]))
                a["ms"] += kt["ms"]
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = (g.timings()["kernel_launches"] - launches0) // args.steps
    counts = g.counts()
    tile_counts = h0.counts()
    ms = sum(times) / len(times)

    for _ in range(2):
        step_e2e()
    barrier()
    te = []
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        te.append((time.perf_counter() - t0) * 1e3)
    barrier()
    ms_e2e = sum(te) / len(te)
    same = bool(torch.equal(host_lab.cuda(), dev_lab))

    t = torch.tensor([ms, ms_e2e, float(launches)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    tl = torch.tensor([float(launches), 1.0 if same else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    launches_all = int(tl[0])
    same_all = int(tl[1]) == world

    # ---- parity: the slab labels against ONE single-GPU run of the whole scene (rank 0); scenes too large for one GPU:
    #      the size-independent properties of canonical labels ----
    parity = None
    if not args.no_verify:
        sizes = [ppp * (nparts // world)] * world
        gathered = [torch.empty(sz, dtype=torch.int32, device="cuda") for sz in sizes] if rank == 0 else None
        dist.gather(dev_lab, gathered, dst=0)
        if rank == 0:
            try:
                got_dev = torch.cat(gathered)
                del gathered
                if n_total <= 250_000_000:
                    whole = np.concatenate([pts] + [make_part(cfg, CONFIGS, j, ppp, extent) for j in range(len(my_parts), nparts)], axis=0)
                    hs = capi.Handle(mode=0, device=local, stream=stream.cuda_stream)
                    hs.set_points(whole)
                    ref = hs.run(params)
                    c1 = hs.counts()
                    hs.close()
                    got = got_dev.cpu().numpy()
                    parity = {"labels_equal_single_gpu": bool(np.array_equal(got, ref)), "points": int(whole.shape[0]),
                              "clusters_single_gpu": int(c1["n_clusters_exported"]), "clusters_slabs": int(counts["n_clusters_exported"])}
                else:
                    lab = got_dev.long()
                    idx = torch.arange(lab.numel(), device="cuda")
                    has = lab >= 0
                    seeds = lab[has]
                    parity = {"single_gpu_comparison": "skipped: the scene does not fit one GPU (the same split is compared with one GPU on "
                                                       "urban100m and in tests/test_slabs_gpu.py)",
                              "points": int(lab.numel()), "labelled_points": int(has.sum()),
                              "label_is_min_index_of_its_cluster": bool((seeds <= idx[has]).all()) and bool((lab[seeds] == seeds).all()),
                              "distinct_labels": int(torch.unique(seeds).numel()), "clusters_slabs": int(counts["n_clusters_exported"])}
            except Exception as e:  # noqa: BLE001
                parity = {"error": str(e)[:300]}
    g.close()
    if rank != 0:
        return None
    peak, peak_kind = measured_peak()
    kernels = []
    for name, a in kern_acc.items():
        t_ms = a["ms"] / args.steps
        gbs = a["alg_bytes"] / t_ms / 1e6 if t_ms > 0 else 0.0
        kernels.append({"kernel": name, "ms": round(t_ms, 4), "launches": a["launches"], "alg_bytes": a["alg_bytes"],
                        "GBps": round(gbs, 1), "frac_of_hbm": round(gbs / peak, 4), "share_of_step": round(t_ms / ms, 4)})
    dom = max(kernels, key=lambda kk: kk["ms"])
    strong = CONFIGS[cfg]["scene"] == "urban"
    return {
        "metric": "points/sec segmented end-to-end", "value": n_total / (ms / 1e3), "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CONFIGS[cfg]['what']}, ONE scene of {n_total} points ({nparts} parts of {ppp} on {extent:.1f} m), "
                               f"{n} points per GPU before routing", "name": cfg,
                   "multi_gpu_mode": "slabs: one scene, spatial slabs along the longest axis + halo voxels, global PCL origin, "
                                     "closest-check rounds and cross-slab union-find over NCCL; labels bit-equal to one GPU",
                   "points_total": n_total, "points_per_gpu": n,
                   "l2": "512 MB buffer written between timed steps (L2 flush)",
                   "halo_voxel_layers": counts["halo"], "slab_axis": counts["axis"], "cuts": counts["cuts"],
                   "tile_points_all_ranks": counts["n_tile_points"], "halo_overhead": counts["n_tile_points"] / max(1, n_total) - 1.0,
                   "tile_voxels_all_ranks": counts["n_tile_voxels"], "cross_slab_pairs": counts["n_cross_pairs"],
                   "origin_rounds": counts["origin_rounds"], "closest_rounds": counts["closest_rounds"],
                   "clusters": counts["n_clusters_exported"], "octree_depth": counts["octree_depth"],
                   "rank0_tile": {"voxels": tile_counts["n_voxels"], "used": tile_counts["n_used"], "adjacency_entries": tile_counts["n_adjacency"]}},
        "e2e": {"value": n_total / (ms_e2e / 1e3), "unit": "points/s", "h2d_bytes_per_step": 12 * n_total, "d2h_bytes_per_step": 4 * n_total,
                "ms_per_step": ms_e2e, "labels_equal_resident_path": same_all},
        "gpu_launches": launches_all,
        "parity": parity,
        "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                     "frac": dom["GBps"] / peak, "traffic": None, "peak_kind": peak_kind, "algorithmic_bytes": dom["alg_bytes"], "ms": dom["ms"],
                     "note": "rank 0's tile: the kernel group with the largest share of its step; every group is in `kernels`"},
        "kernels": kernels,
        "group_stages_ms": {k[:-3]: round(v / args.steps, 4) for k, v in gt_acc.items() if k.endswith("_ms")},
        "clocks": clk.summary(),
        "wall_s_timed_region": t_wall,
    }
