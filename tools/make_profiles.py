"""Turn the raw ncu / bench outputs of a gpurun call into the committed summaries under profiles/ (dev tool).
   python tools/make_profiles.py r01 gpurun_out/launches_r1_final.csv gpurun_out/bench_r1_final.json [gpurun_out/prof.ncu-rep]"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, bench_json = sys.argv[1:4]
rep = sys.argv[4] if len(sys.argv) > 4 else None

rows = list(csv.reader(open(launches_csv)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
per = collections.defaultdict(dict)
name = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
    name[r[ii]] = r[ki].split("(")[0].replace("vgs::", "").split("<")[0]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, m in per.items():
    a = agg[name[k]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0) / 1e6
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
out = [f"# {tag} — ncu launch list of the final build (one pipeline pass, 10 M-point construction site)", "",
       "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv",
       "python tools/profile_run.py 10000000 1` (one VGS pass through the C ABI, host points in, host labels out).",
       "Per-launch times are cold-cache and serialised: compare SHARES with `bench.py`'s CUDA-event stage timers.", "",
       "| kernel | launches | total ms | share | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / total:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} |")
out += ["", f"Total {total:.2f} ms over {sum(a[0] for a in agg.values())} launches.", ""]
st = {k: v["ms"] for k, v in bench["stages"].items()}
out += [f"Same build, `python bench.py --steps {bench['steps']} --warmup {bench['warmup']}` (`profiles/{tag}_bench_10M_final.json`): "
        f"{bench['ms_per_step']:.2f} ms per step ({bench['value']:.3e} points/s device-resident, {bench['e2e']['value']:.3e} points/s "
        f"through host buffers, {bench['gpu_launches']} launches in {bench['steps']} steps), stage timers (ms): {json.dumps(st)}.",
        f"Share check: graph stage (k_pair_cache + k_bin_classes + k_rs_* of the class lists + k_local_graph_*) = "
        f"{100 * st['graph'] / sum(st.values()):.1f}% of the CUDA-event stage sum; the same kernels are "
        f"{100 * sum(a[1] for k, a in agg.items() if k in ('k_local_graph_warp', 'k_local_graph2', 'k_pair_cache', 'k_pair_cache_bm', 'k_bitmap_set', 'k_bin_classes', 'k_class_init')) / total:.1f}% "
        "of the ncu total."]
open(os.path.join(ROOT, "profiles", f"{tag}_launches_10M_final.md"), "w").write("\n".join(out) + "\n")
dom = ("k_local_graph_warp", "k_local_graph2", "k_pair_cache", "k_pair_cache_bm", "k_bitmap_set", "k_bin_classes", "k_class_init")
json.dump({"kernel": "stage 4+5a: k_pair_cache_bm + k_bin_classes + k_local_graph_warp / k_local_graph2 (all launches of one pipeline pass, 10 M-point construction site)",
           "launches": sum(agg[k][0] for k in dom if k in agg),
           "dram_bytes_read": sum(agg[k][2] for k in dom if k in agg), "dram_bytes_write": sum(agg[k][3] for k in dom if k in agg),
           "gpu_time_ms_ncu": sum(agg[k][1] for k in dom if k in agg),
           "per_kernel": {k: {"launches": agg[k][0], "ms": agg[k][1], "read": agg[k][2], "write": agg[k][3]} for k in dom if k in agg},
           "source": f"profiles/{tag}_launches_10M_final.md (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none)"},
          open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
print("\n".join(out[:14]))

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    names, units, vals = rr[0], rr[1], rr[2]
    d = {n: (v, u) for n, u, v in zip(names, units, vals)}
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio"]
    lines = ["| metric | value |", "|---|---|"]
    for w in want:
        if w in d:
            lines.append(f"| {w} | {d[w][0]} {d[w][1]} |")
    stalls = {n.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(v.replace(",", ""))
              for n, v in zip(names, vals) if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and v}
    top = ", ".join(f"{k} {v:.2f}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v > 0.3)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_raw_table.md"), "w").write("\n".join(lines) + "\n\nWarp stall reasons (warps stalled per issue slot, > 0.3): " + top + "\n")
    print("\n".join(lines), "\n", top)
