#!/usr/bin/env python
"""bench.py — points/sec segmented end-to-end (BASELINE.json metric) on N B200s of one node.

A "step" = one full VGS segmentation (voxelise -> features -> adjacency -> local graphs ->
mutual filter -> closest check -> components -> per-point labels) of one synthetic
construction-site scene (BASELINE.json configs[2]: 10 M points, Task_File_VGS.txt parameters).

  value : whole-job points/s with the point cloud already resident in HBM (device pointer in,
          device labels out), CUDA-event timed per step, L2 flushed between steps.
  e2e   : the same through the host-buffer C-ABI call (pinned host xyz in, host labels out,
          H2D + D2H inside the timed region).
  roofline : the dominant stage (local graph kernels), algorithmic bytes / event time / measured HBM peak.
  cpu_baseline : the CPU oracle (single thread, like the reference) on a bounded sample.

`--impl reference` times the CPU restatement of the reference (oracle, glibc libm) instead.
Multi-GPU (torchrun): every rank segments its own tile of the site grid (weak scaling), no
data-path collective inside the timed region except the barrier; max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VGS_PARAMS = dict(voxel_size=0.15, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2, sig_c=0.2, sig_w=2.0,
                  cut_thred=0.3, points_min=10, adjacency_min=3, voxels_min=3)
HBM_FALLBACK_GBS = 6650.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region, through NVML in-process (one
    `nvidia-smi -lms` child as fallback) so that the sampling itself does not perturb the GPU."""
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index = index
        self.rows = []          # (sm_mhz, sm_max_mhz, reasons_bitmask)
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)
        self.child = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(hnd, pynvml.NVML_CLOCK_SM)
            while not self.stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(hnd)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hnd)
                self.rows.append((float(sm), float(mx), int(rs)))
                self.stop.wait(0.02)
            return
        except Exception:
            pass
        try:   # fallback: ONE long-running nvidia-smi child, as in the profiling recipe
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            self.child = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                           "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.child.stdout:
                f = [x.strip() for x in line.split(",")]
                try:
                    self.rows.append((float(f[0]), float(f[1]), int(f[2], 16)))
                except Exception:
                    pass
                if self.stop.is_set():
                    break
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.child is not None:
            try:
                self.child.terminate()
            except Exception:
                pass
        self.th.join(timeout=3)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(n for n, bit in self.BITS.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in self.rows), default=None),
                "reasons": reasons, "samples": len(self.rows)}


def make_scene(n_points, rank):
    from vgs_svgs_segmentation_b200 import scenes
    # each rank = one 70 m tile of a site grid (tile origin shifted by 80 m in x), own seed
    extent = 70.0 * (n_points / 10_000_000) ** 0.5
    return scenes.construction_site(n_points, seed=1 + rank, extent=extent, offset=(80.0 * rank, 0.0, 0.0))


def cpu_baseline(sample_points, math=0, threads=1):
    """threads = 1: as the reference runs (it has no threading); threads > 1: the per-unit local-graph loop of the oracle
    (the dominant cost) over OpenMP threads, same results"""
    from oracle import oracle
    pts = make_scene(sample_points, 0)
    cores = oracle.set_threads(threads)
    t0 = time.perf_counter()
    r = oracle.run(pts, math=math)
    dt = time.perf_counter() - t0
    oracle.set_threads(1)
    how = "single thread, as the reference runs" if cores == 1 else f"{cores} OpenMP threads on the per-unit local-graph loop"
    return {"value": sample_points / dt, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": f"construction_site {sample_points} points (same density as the workload), CPU oracle, {how}, "
                      f"{dt:.2f} s, {r.stats['pair_evals']} pair evaluations"}, pts, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    per = []
    info = None
    threads = os.cpu_count() or 1       # "all the host threads it can use"; the reference itself is single-threaded
    for i in range(warm + steps):
        info, _, _ = cpu_baseline(args.ref_points, math=0, threads=threads)
        if i >= warm:
            per.append(args.ref_points / info["value"])
    ms = 1e3 * sum(per) / len(per)
    val = args.ref_points / (ms / 1e3)
    info["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": "points/sec segmented end-to-end", "value": val, "unit": "points/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"VGS, synthetic construction-site scene {args.points} points per GPU (BASELINE.json configs[2]), "
                               "Task_File_VGS.txt parameters (voxel 0.15, graph 0.5, sigma 0.2 x5, sig_w 2, cut 0.3, "
                               "points_min 10, adjacency_min 3, voxels_min 3)",
                   "sample": f"each step = the CPU oracle on a bounded {args.ref_points}-point sample of that scene (same density)"},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference cannot be compiled (needs PCL 1.8.1, and voxel_segmentation.h:2279 is undefined); "
                "this is the CPU oracle restatement; the reference is single-threaded, here its per-unit loop runs on all host "
                "threads (cpu_baseline.cores); the single-thread figure is the cpu_baseline of the default arm"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--ref-points", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample", type=int, default=400_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mode", default="tiles", choices=["tiles", "partitioned"],
                    help="N>1: 'tiles' = one independent site tile per rank (weak scaling, no exchange); "
                         "'partitioned' = ONE scene, local-graph stage split over ranks + NCCL exchange (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build(oracle=True, quiet=True)
    from vgs_svgs_segmentation_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    partitioned = args.mode == "partitioned" and world > 1
    pts = make_scene(args.points, 0 if partitioned else rank)
    n = pts.shape[0]
    host_xyz = torch.from_numpy(pts).pin_memory()
    host_lab = torch.empty(n, dtype=torch.int32).pin_memory()
    dev_xyz = host_xyz.cuda(non_blocking=False)
    dev_lab = torch.empty(n, dtype=torch.int32, device="cuda")
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 512 MB > 126 MB L2
    stream = torch.cuda.current_stream()
    h = capi.Handle(mode=capi.VGS_MODE_VGS, device=local, stream=stream.cuda_stream)
    params = capi.make_params(**VGS_PARAMS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from vgs_svgs_segmentation_b200.multigpu import segment_partitioned

    def step_resident():
        h.set_points_device(dev_xyz.data_ptr(), n, 12)
        if partitioned:
            segment_partitioned(h, params, rank, world, dev_lab.data_ptr(), on_device=True)
        else:
            h.run(params, dev_lab.data_ptr(), on_device=True)

    def step_e2e():
        h.set_points_host_ptr(host_xyz.data_ptr(), n, 12)
        if partitioned:
            segment_partitioned(h, params, rank, world, host_lab.numpy(), on_device=False)
        else:
            h.run(params, host_lab.numpy(), on_device=False)
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_resident()
    launches0 = h.timings()["kernel_launches"]
    stage_acc = {}
    times = []
    barrier()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_resident()
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
            for k, v in h.timings().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = h.timings()["kernel_launches"] - launches0
    counts = h.counts()
    ms = sum(times) / len(times)

    # ---- end-to-end (host buffers through the C ABI) ----
    for _ in range(2):
        step_e2e()
    barrier()
    te = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_e2e()
        te.append((time.perf_counter() - t0) * 1e3)
    barrier()
    ms_e2e = sum(te) / len(te)
    # the e2e labels must equal the resident-path labels
    same = bool(torch.equal(host_lab.cuda(), dev_lab))

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        tot = torch.tensor([n], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        total_points = n if partitioned else int(tot[0])
    else:
        total_points = n

    if rank == 0:
        peak, peak_kind = measured_peak()
        V, E, N = counts["n_units"], counts["n_adjacency"], counts["n_points"]
        per_stage = {k: stage_acc[k] / args.steps for k in stage_acc if k.endswith("_ms")}
        # algorithmic bytes per stage (SURVEY.md §8d / DESIGN.md §kernels)
        alg = {"origin_ms": 12 * N, "voxelize_ms": 20 * N + 16 * V, "features_ms": 16 * N + 64 * V,
               "adjacency_ms": 32 * V + 4 * E, "graph_ms": 8 * E + 64 * V, "mutual_ms": 12 * E,
               "components_ms": 8 * E + 8 * V, "labels_ms": 8 * N + 4 * V}
        stages = {}
        for k, b in alg.items():
            t_ms = per_stage.get(k, 0.0)
            if t_ms > 0:
                stages[k[:-3]] = {"ms": round(t_ms, 4), "alg_bytes": int(b), "GBps": round(b / t_ms / 1e6, 2),
                                  "frac_of_hbm": round(b / t_ms / 1e6 / peak, 4)}
        traffic = None
        tf = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tf) and n == 10_000_000 and not partitioned:
            try:
                t_ = json.load(open(tf))
                traffic = float(t_["dram_bytes_read"]) + float(t_["dram_bytes_write"])
            except Exception:
                traffic = None
        dom = max(per_stage, key=lambda k: per_stage[k] if k in alg else -1)
        dom_b = alg[dom]
        achieved = dom_b / per_stage[dom] / 1e6
        out = {
            "metric": "points/sec segmented end-to-end", "value": total_points / (ms / 1e3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if partitioned else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"VGS, ONE synthetic construction-site scene of {n} points, local-graph stage partitioned over "
                                    f"{world} GPUs + NCCL exchange of connect lists, " if partitioned else
                                    f"VGS, synthetic construction-site scene {n} points per GPU (BASELINE.json configs[2]), ") +
                                   "Task_File_VGS.txt parameters (voxel 0.15, graph 0.5, sigma 0.2 x5, sig_w 2, cut 0.3, "
                                   "points_min 10, adjacency_min 3, voxels_min 3)",
                       "multi_gpu_mode": args.mode if world > 1 else "single", "points_per_gpu": n, "l2": "512 MB buffer written between timed steps (L2 flush); per-step working set > 1 GB",
                       "tiles": world, "voxels": V, "used_voxels": counts["n_used"], "adjacency_entries": E,
                       "pair_weights": counts["n_pairs"], "clusters": counts["n_clusters_exported"]},
            "e2e": {"value": total_points / (ms_e2e / 1e3), "unit": "points/s", "h2d_bytes_per_step": 12 * n,
                    "d2h_bytes_per_step": 4 * n, "ms_per_step": ms_e2e, "labels_equal_resident_path": same},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pair_cache_bm + k_bin_classes + k_local_graph_warp (stage 4+5a, all size classes)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_kind": peak_kind, "algorithmic_bytes": int(dom_b),
                         "algorithmic_bytes_no_reuse": int(72 * E + 64 * V),
                         "sm_issue_active_pct_ncu": 60.8,
                         "note": "achieved = algorithmic bytes (8*E + 64*V, SURVEY.md 8d / DESIGN.md section 4) / CUDA-event time of the "
                                 "stage; SURVEY.md 8d classifies this stage as compute/latency bound (sum of n^2 pair work), not HBM bound: "
                                 "ncu shows 61 % issue-slot utilisation and 8 % DRAM throughput for k_local_graph_warp "
                                 "(profiles/r01_ncu_local_graph_warp_final.md); algorithmic_bytes_no_reuse = 64*E record gathers + 8*E if no "
                                 "record were reused; traffic = ncu dram read+write bytes of the stage's kernels per pass on this workload "
                                 "(profiles/r01_traffic.json): the pair table rows (8.8 KB, sparse) and the bin-ordered entry scratch"},
            "stages": stages,
            "clocks": clk.summary(),
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu:
            info, _, _ = cpu_baseline(args.cpu_sample, math=0)
            out["cpu_baseline"] = info
        print(json.dumps(out))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
