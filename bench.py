#!/usr/bin/env python
"""bench.py — points/sec segmented end-to-end (BASELINE.json metric) on N B200s of one node.

A "step" = one full segmentation of one synthetic scene through the C ABI of libvgs_b200.so:
  VGS : voxelise -> features -> adjacency -> weight rows -> local graphs -> mutual filter -> closest check -> components
        -> per-point labels (Task_File_VGS.txt parameters)
  SVGS: voxelise -> VCCS supervoxels (CUDA generator) -> the same pipeline on supervoxels (Task_File_SVGS.txt parameters)

  --config site10m (default) : BASELINE.json configs[2], 10 M-point construction site, VGS           [the headline line]
           town2m            : configs[0] stand-in (Town_Test.pcd is not distributed), 2 M points, VGS
           town2m_svgs       : configs[1] stand-in, 2 M points, SVGS
           urban100m         : configs[3], 100 M-point urban scene, VGS (one GPU: fits in 180 GB)
           city1b            : configs[4], 1 B-point city scene, VGS, only with --gpus 8 (slab split; 125 M points per GPU)

  value    : whole-job points/s, point cloud already resident in HBM (device pointer in, device labels out), CUDA events
             around every step on the launching stream, a 512 MB buffer written between steps (L2 flush).
  e2e      : the same through the host-buffer call (pinned host xyz in, host labels out; H2D + D2H inside the timed region).
  roofline : the kernel group with the largest share of the step; `kernels` lists every group (CUDA-event time measured
             live inside the library, algorithmic bytes from vgs_kernel_timings, fraction of the measured HBM peak).
  cpu_baseline : the CPU oracle, single thread (as the reference runs), on a bounded sample of the same workload.

`--impl reference` times the CPU restatement of the reference (the oracle; the reference itself cannot be compiled here)
on the SAME scene with all host threads, and checks its labels against one run of the CUDA path when a GPU is present.

N > 1 (torchrun, one rank per GPU): `--mode slabs` (default for VGS) segments ONE scene split into spatial slabs with halo
voxels, global PCL origin, closest-check rounds and cross-slab component merge over NCCL (vgs_group_run), and checks the
labels against one single-GPU run of the whole scene.  site10m / town2m: the scene has N parts of the config's size on
sqrt(N) times the extent (weak scaling, N = 1 is the single-GPU workload); urban100m: the same 100 M-point scene for every N
(strong scaling).  `--mode replicas` runs one independent scene per rank (no exchange).
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
SVGS_PARAMS = dict(voxel_size=0.05, graph_size=0.5, sig_p=0.2, sig_n=0.2, sig_o=0.2, sig_e=0.2, sig_c=0.75, sig_w=1.0,
                   cut_thred=0.5, points_min=10, adjacency_min=3, voxels_min=3)
VCCS = dict(seed_resolution=0.25, color_importance=0.0, spatial_importance=0.25, normal_importance=0.75, refine_iterations=5)
HBM_FALLBACK_GBS = 6650.0

CONFIGS = {
    "site10m": dict(scene="construction_site", points=10_000_000, mode=0,
                    what="VGS, synthetic construction-site scene (BASELINE.json configs[2]), Task_File_VGS.txt parameters"),
    "town2m": dict(scene="town", points=2_000_000, mode=0,
                   what="VGS, synthetic town scene standing in for Town_Test.pcd (configs[0]), Task_File_VGS.txt parameters"),
    "town2m_svgs": dict(scene="town", points=2_000_000, mode=1,
                        what="SVGS, synthetic town scene standing in for Town_Test.pcd (configs[1]), Task_File_SVGS.txt parameters, "
                             "supervoxels by the CUDA VCCS generator inside the step"),
    "urban100m": dict(scene="urban", points=100_000_000, mode=0,
                      what="VGS, synthetic Semantic3D-scale urban scene (configs[3]), Task_File_VGS.txt parameters"),
    "city1b": dict(scene="urban", points=1_000_000_000, mode=0,
                   what="VGS, synthetic 1 B-point tiled city scene (configs[4]: 80 urban tiles of 12.5 M points on one 1012 m ground), "
                        "Task_File_VGS.txt parameters; needs the slab split over several GPUs"),
}


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
                self.stop.wait(0.005)
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


def make_scene(cfg, n_points=None, tile=0):
    """the config's scene at n_points (default: its full size); extents scale with sqrt(points) so the density stays that of
    the full scene; tile > 0 shifts a replica (its own seed) by 80 m in x.  The 100 M-point urban scene is defined as 8 parts of
    12.5 M points (seeds 2..9) on the same 320 m ground, so that every GPU count segments the same cloud (slabs.scene_parts)."""
    from vgs_svgs_segmentation_b200 import scenes
    c = CONFIGS[cfg]
    n = int(n_points or c["points"])
    if c["scene"] == "construction_site":
        return scenes.construction_site(n, seed=1 + tile, extent=70.0 * (n / 10_000_000) ** 0.5, offset=(80.0 * tile, 0.0, 0.0))
    if c["scene"] == "town":
        return scenes.town(n, seed=20170610 + tile, extent=60.0 * (n / 2_000_000) ** 0.5, offset=(80.0 * tile, 0.0, 0.0))
    import numpy as np
    ext = 320.0 * (n / 100_000_000) ** 0.5
    parts = max(1, n // 12_500_000)
    return np.concatenate([scenes.urban(n // parts, seed=2 + j + 1000 * tile, extent=ext, offset=(4000.0 * tile, 0.0, 0.0)) for j in range(parts)], axis=0)


def params_of(cfg):
    return dict(SVGS_PARAMS if CONFIGS[cfg]["mode"] == 1 else VGS_PARAMS)


def oracle_run(cfg, pts, threads, math):
    """the CPU restatement on `pts`; SVGS: supervoxel labels by the oracle's own VCCS restatement (synchronous schedule)"""
    from oracle import oracle
    pd = params_of(cfg)
    cores = oracle.set_threads(threads)
    t0 = time.perf_counter()
    try:
        if CONFIGS[cfg]["mode"] == 1:
            v = oracle.vccs(pts, voxel_res=pd["voxel_size"], seed_res=VCCS["seed_resolution"], color_importance=VCCS["color_importance"],
                            spatial_importance=VCCS["spatial_importance"], normal_importance=VCCS["normal_importance"],
                            refine_iterations=VCCS["refine_iterations"], schedule=1)
            r = oracle.run(pts, labels=v.point_label, max_label=v.max_label, mode=1, math=math, **pd)
        else:
            r = oracle.run(pts, math=math, **pd)
    finally:
        oracle.set_threads(1)
    return r, time.perf_counter() - t0, cores


def cpu_baseline(cfg, sample_points):
    """single thread, as the reference runs (it has no threading), on a bounded sample of the workload"""
    pts = make_scene(cfg, sample_points)
    r, dt, cores = oracle_run(cfg, pts, 1, 0)
    return {"value": sample_points / dt, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": f"{CONFIGS[cfg]['scene']} scene cut down to {sample_points} points at the density of the workload, CPU oracle "
                      f"(glibc libm), single thread as the reference runs, {dt:.2f} s, {r.stats['pair_evals']} pair evaluations"}


def dropin_e2e(cfg, pts, pd, runs=3):
    """e2e through the reference-facing C++ class (include/vgs_dropin/voxel_segmentation.h) driven by the reference's call sequence
    (test:51-76): pageable pcl::PointCloud (16-byte points) in, drawColorMapofPointsinClusters + getClusterIdx() (vector<vector<int>>)
    out.  A separate process (tests/cpp/dropin_vgs.cpp, compiled here with g++); VGS only."""
    import tempfile
    exe = os.path.join(ROOT, "tests", "_build", "dropin_vgs")
    try:
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        src = os.path.join(ROOT, "tests", "cpp", "dropin_vgs.cpp")
        deps = [src] + [os.path.join(ROOT, "include", "vgs_dropin", f) for f in os.listdir(os.path.join(ROOT, "include", "vgs_dropin"))]
        if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
            subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "include"), "-o", exe, src,
                            "-L" + os.path.join(ROOT, "vgs_svgs_segmentation_b200"), "-lvgs_b200",
                            "-Wl,-rpath," + os.path.join(ROOT, "vgs_svgs_segmentation_b200")], check=True, capture_output=True)
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "x.f32")
            pts.tofile(f)
            args = [exe, f, str(pts.shape[0]), os.path.join(td, "l.i32")] + [str(pd[k]) for k in (
                "voxel_size", "graph_size", "sig_p", "sig_n", "sig_o", "sig_e", "sig_c", "sig_w", "cut_thred", "points_min", "adjacency_min", "voxels_min")]
            r = subprocess.run(args, capture_output=True, text=True, timeout=900, env=dict(os.environ, VGS_DROPIN_REPEAT=str(runs + 1)))
            if r.returncode != 0:
                return {"error": (r.stderr or r.stdout)[-300:]}
            ms = [float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("dropin_ms")]
            phases = [[float(x) for x in l.split()[1:]] for l in r.stdout.splitlines() if l.startswith("dropin_phases")]
            import numpy as np
            lab = np.fromfile(os.path.join(td, "l.i32"), np.int32)
        v = sum(ms) / len(ms)
        names = ("voxelise_incl_h2d", "centres", "features", "adjacency", "segment", "clusters_and_coloured_cloud", "getClusterIdx", "destruction")
        phase_ms = {k: round(sum(p[i] for p in phases) / len(phases), 3) for i, k in enumerate(names)} if phases else None
        return {"value": pts.shape[0] / (v / 1e3), "unit": "points/s", "ms_per_run": v, "runs": len(ms), "labels": lab, "phases_ms": phase_ms,
                "what": "pcl::VoxelBasedSegmentation drop-in class, the reference's call sequence (test:51-76) incl. getVoxelCenters, "
                        "drawColorMapofPointsinClusters (an XYZRGB cloud of every clustered point) and getClusterIdx() (vector<vector<int>>); "
                        "pageable host cloud, a fresh object per run, constructed and destroyed inside the timed region (the library parks its device "
                        "working set between objects: vgs_acquire / vgs_release)"}
    except Exception as e:   # noqa: BLE001
        return {"error": str(e)[:300]}


def gpu_step(h, capi, cfg, pd, labels_ptr_or_array, on_device):
    """one full segmentation on handle h (points already set) through the C-ABI calls"""
    if CONFIGS[cfg]["mode"] == 0:
        h.run(capi.make_params(**pd), labels_ptr_or_array, on_device=on_device)
        return
    h.voxelize(pd["voxel_size"])
    h.make_supervoxels_vccs(**VCCS)
    h.compute_features(pd["points_min"])
    h.find_adjacency(pd["graph_size"])
    h.segment(capi.Sigmas(pd["sig_p"], pd["sig_n"], pd["sig_o"], pd["sig_e"], pd["sig_c"], pd["sig_w"]), pd["cut_thred"], pd["adjacency_min"])
    ptr = labels_ptr_or_array if isinstance(labels_ptr_or_array, int) else labels_ptr_or_array.ctypes.data
    h._ck(h.L.vgs_get_point_labels(h.h, pd["voxels_min"], capi.C.c_void_p(ptr), 1 if on_device else 0))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import __graft_entry__ as g
    g.build(oracle=True, quiet=True)
    cfg = args.config
    threads = os.cpu_count() or 1          # "all the host threads it can use"; the reference itself is single-threaded
    pts = make_scene(cfg, args.ref_points or None)
    n = pts.shape[0]
    per = []
    r = None
    warm = min(args.warmup, 1)             # every step is tens of seconds of CPU work: one warm-up pass is plenty
    for i in range(warm + args.steps):
        r, dt, cores = oracle_run(cfg, pts, threads, 1)
        if i >= warm:
            per.append(dt)
    ms = 1e3 * sum(per) / len(per)
    val = n / (ms / 1e3)
    parity = None
    try:      # the CPU labels against one run of the CUDA path on the same scene (when a GPU is here)
        import torch
        if torch.cuda.is_available() and CONFIGS[cfg]["mode"] == 0:
            from vgs_svgs_segmentation_b200 import capi
            h = capi.Handle(mode=0, device=0)
            h.set_points(pts)
            lab = h.run(capi.make_params(**params_of(cfg)))
            h.close()
            parity = {"labels_equal": bool(np.array_equal(lab, r.point_label)), "points": int(n),
                      "near_threshold_decisions": int(r.stats["near_threshold"])}
    except Exception as e:   # noqa: BLE001
        parity = {"error": str(e)[:200]}
    info = {"value": val, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": f"the full {CONFIGS[cfg]['scene']} scene of the config ({n} points) per step, CPU oracle with correctly rounded libm "
                      f"(the definition the CUDA path is compared with), per-unit local-graph loop on {cores} OpenMP threads"}
    print(json.dumps({
        "impl": "reference", "metric": "points/sec segmented end-to-end", "value": val, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CONFIGS[cfg]['what']}, {n} points", "name": cfg, "same_scene_as_default_arm": args.ref_points in (0, None)},
        "cpu_baseline": info, "parity": parity,
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference cannot be compiled (needs PCL 1.8.1, and voxel_segmentation.h:2279 is undefined); this is the CPU oracle "
                "restatement; the reference is single-threaded, here its dominant per-unit loop runs on all host threads (cpu_baseline.cores); "
                "the single-thread figure is the cpu_baseline of the default arm"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--points", type=int, default=0, help="override the number of points of the config's scene (density kept)")
    ap.add_argument("--ref-points", type=int, default=0, help="--impl reference: 0 = the full scene of the config")
    ap.add_argument("--cpu-sample", type=int, default=600_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the e2e_dropin figure (the C++ drop-in class driven like the reference's test)")
    ap.add_argument("--no-verify", action="store_true", help="N>1 slabs: skip the comparison with one single-GPU run of the whole scene")
    ap.add_argument("--mode", default=None, choices=["slabs", "replicas"],
                    help="N>1: 'slabs' = ONE scene split into spatial slabs with halo voxels + cross-slab merge over NCCL (strong scaling); "
                         "'replicas' = one independent tile per rank (weak scaling, no exchange)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config is None:
        args.config = "site10m"
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build(oracle=True, quiet=True)
    from vgs_svgs_segmentation_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = args.config
    cmode = CONFIGS[cfg]["mode"]
    pd = params_of(cfg)
    mode = args.mode or ("slabs" if cmode == 0 else "replicas")
    slabs = mode == "slabs" and world > 1 and cmode == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 512 MB > 126 MB L2
    if slabs:
        from vgs_svgs_segmentation_b200 import slabs as slabmod
        out = slabmod.bench(args, cfg, pd, rank, world, local, flush, barrier, ClockSampler, measured_peak, CONFIGS)
        if rank == 0:
            if not args.no_cpu:
                out["cpu_baseline"] = cpu_baseline(cfg, args.cpu_sample)
            print(json.dumps(out))
        dist.destroy_process_group()
        return

    pts = make_scene(cfg, args.points or None, tile=rank)
    n = pts.shape[0]
    host_xyz = torch.from_numpy(pts).pin_memory()
    host_lab = torch.empty(n, dtype=torch.int32).pin_memory()
    dev_xyz = host_xyz.cuda(non_blocking=False)
    dev_lab = torch.empty(n, dtype=torch.int32, device="cuda")
    h = capi.Handle(mode=cmode, device=local, stream=stream.cuda_stream)

    def step_resident():
        h.set_points_device(dev_xyz.data_ptr(), n, 12)
        gpu_step(h, capi, cfg, pd, dev_lab.data_ptr(), True)

    def step_e2e():
        h.set_points_host_ptr(host_xyz.data_ptr(), n, 12)
        gpu_step(h, capi, cfg, pd, host_lab.numpy(), False)
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_resident()
    launches0 = h.timings()["kernel_launches"]
    stage_acc, kern_acc, times = {}, {}, []
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
            for kt in h.kernel_timings():
                a = kern_acc.setdefault(kt["name"], dict(ms=0.0, launches=kt["launches"], alg_bytes=kt["alg_bytes"]))
                a["ms"] += kt["ms"]
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = (h.timings()["kernel_launches"] - launches0) // args.steps
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
    same = bool(torch.equal(host_lab.cuda(), dev_lab))     # the e2e labels must equal the resident-path labels

    e2e_dropin = None
    if world == 1 and cmode == 0 and not args.no_dropin:
        e2e_dropin = dropin_e2e(cfg, pts, pd)
        if "labels" in e2e_dropin:
            e2e_dropin["labels_equal_resident_path"] = bool(np.array_equal(e2e_dropin.pop("labels"), dev_lab.cpu().numpy()))

    if world > 1:     # max over ranks
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        tot = torch.tensor([n], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        total_points = int(tot[0])
    else:
        total_points = n

    if rank == 0:
        peak, peak_kind = measured_peak()
        per_stage = {k: stage_acc[k] / args.steps for k in stage_acc if k.endswith("_ms")}
        kernels = []
        for name, a in kern_acc.items():
            t_ms = a["ms"] / args.steps
            gbs = a["alg_bytes"] / t_ms / 1e6 if t_ms > 0 else 0.0
            kernels.append({"kernel": name, "ms": round(t_ms, 4), "launches": a["launches"], "alg_bytes": a["alg_bytes"],
                            "GBps": round(gbs, 1), "frac_of_hbm": round(gbs / peak, 4), "share_of_step": round(t_ms / ms, 4)})
        dom = max(kernels, key=lambda kk: kk["ms"])
        out = {
            "metric": "points/sec segmented end-to-end", "value": total_points / (ms / 1e3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{CONFIGS[cfg]['what']}, {n} points per GPU", "name": cfg,
                       "multi_gpu_mode": "replicas (one independent tile per rank, no exchange)" if world > 1 else "single",
                       "points_per_gpu": n, "l2": "512 MB buffer written between timed steps (L2 flush); per-step working set > 1 GB",
                       "voxels": counts["n_voxels"], "units": counts["n_units"], "used_units": counts["n_used"],
                       "adjacency_entries": counts["n_adjacency"], "pair_weights_of_the_reference": counts["n_pairs"],
                       "clusters": counts["n_clusters_exported"], "octree_depth": counts["octree_depth"]},
            "e2e": {"value": total_points / (ms_e2e / 1e3), "unit": "points/s", "h2d_bytes_per_step": 12 * n,
                    "d2h_bytes_per_step": 4 * n, "ms_per_step": ms_e2e, "labels_equal_resident_path": same},
            "e2e_dropin": e2e_dropin,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                         "frac": dom["GBps"] / peak, "traffic": None, "peak_kind": peak_kind, "algorithmic_bytes": dom["alg_bytes"],
                         "ms": dom["ms"],
                         "note": "the kernel group with the largest share of the step; achieved = algorithmic bytes of the group (DESIGN.md "
                                 "section 4, reported by vgs_kernel_timings) / its CUDA-event time measured in this run; every group is in "
                                 "`kernels`; ncu DRAM traffic per kernel is in profiles/ (not repeated here: it is not measured by this run)"},
            "kernels": kernels,
            "stages_ms": {k[:-3]: round(v, 4) for k, v in per_stage.items() if v > 0},
            "clocks": clk.summary(),
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(cfg, args.cpu_sample)
        print(json.dumps(out))
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
