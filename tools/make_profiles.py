"""Turn the raw ncu / bench outputs of a gpurun call into the committed summaries under profiles/ (dev tool).
   python tools/make_profiles.py r02 gpurun_out/r2_final_launches.csv gpurun_out/r2_final_site10m.json [gpurun_out/r2_final_prof.ncu-rep]
Writes profiles/<tag>_launches_site10m.md (launch list: time + DRAM bytes per kernel against the algorithmic bytes of
bench.py's kernel groups) and profiles/<tag>_ncu_kernels.md (key `--set full` metrics of every captured kernel)."""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, bench_json = sys.argv[1:4]
rep = sys.argv[4] if len(sys.argv) > 4 else None

rows = list(csv.reader(open(launches_csv)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
ui = H.index("Metric Unit")
per = collections.defaultdict(dict)
name = {}
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    per[r[ii]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    name[r[ii]] = re.sub(r"^void ", "", r[ki]).split("(")[0].replace("vgs::", "").split("<")[0]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, m in per.items():
    a = agg[name[k]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
bench = json.loads([l for l in open(bench_json) if l.startswith("{")][-1])
out = [f"# {tag} — ncu launch list of the final build (one pipeline pass, 10 M-point construction site)", "",
       "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv",
       "python tools/profile_run.py 10000000 1` (one VGS pass through the C ABI, host points in, host labels out).",
       "Per-launch times are cold-cache and serialised: compare SHARES with `bench.py`'s CUDA-event kernel timers.", "",
       "| kernel | launches | total ms | share | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / total:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} |")
out += ["", f"Total {total:.2f} ms over {sum(a[0] for a in agg.values())} launches, "
            f"{sum(a[2] + a[3] for a in agg.values()) / 1e9:.2f} GB of DRAM traffic.", ""]
out += [f"Same build, `python bench.py --steps {bench['steps']} --warmup {bench['warmup']}` (`profiles/{tag}_bench_site10m_final.json`): "
        f"**{bench['ms_per_step']:.2f} ms per step** ({bench['value']:.3e} points/s device-resident, {bench['e2e']['value']:.3e} points/s "
        f"through host buffers, {bench['gpu_launches']} launches per step).  Kernel groups (CUDA events inside the library, L2 flushed):", "",
        "| kernel group | ms | share of step | algorithmic MB | GB/s | fraction of the measured HBM peak |", "|---|---:|---:|---:|---:|---:|"]
for k in bench["kernels"]:
    out.append(f"| {k['kernel']} | {k['ms']:.3f} | {100 * k['share_of_step']:.1f}% | {k['alg_bytes'] / 1e6:.0f} | {k['GBps']:.0f} | {k['frac_of_hbm']:.3f} |")
out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_launches_site10m.md"), "w").write("\n".join(out))
print("\n".join(out[:30]))

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    want = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
            ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
            ("launch__occupancy_limit_registers", "CTA/SM limit: registers"), ("launch__occupancy_limit_shared_mem", "CTA/SM limit: smem"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of 64"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
            ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
            ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
            ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
            ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
            ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
            ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
            ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier")]
    kidx = hdr.index("Kernel Name")
    md = [f"# {tag} — `ncu --set full --clock-control none --import-source on` of the top kernels (10 M-point construction site)", "",
          f"Raw report: `{rep}` (scratch).  One launch per row; stalls are warps stalled per issue-active cycle.", ""]
    seen = collections.Counter()
    for r in rr[2:]:
        kn = re.sub(r"^void ", "", r[kidx]).split("(")[0].replace("vgs::", "")
        seen[kn] += 1
        if seen[kn] > 1 and "k_rs_scatter" in kn and seen[kn] != 4:
            continue
        md += [f"## `{kn}`" + (f" (launch {seen[kn]})" if seen[kn] > 1 else ""), "", "| metric | value |", "|---|---:|"]
        for m, label in want:
            if m in hdr:
                i = hdr.index(m)
                md.append(f"| {label} | {r[i]} {units[i]} |")
        md.append("")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.md"), "w").write("\n".join(md))
    print("wrote", f"profiles/{tag}_ncu_kernels.md", sum(seen.values()), "launches")
