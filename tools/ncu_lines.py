"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (compiled with -lineinfo):
   python tools/ncu_lines.py <report.ncu-rep> <kernel name> [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, fname, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "Function Name") and len(r) > 10:
        try:
            ie, isamp, ithr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
            data.append((int(r[ie] or 0), int(r[isamp] or 0), int(r[ithr] or 0), fname, r[0], r[1]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
tots = sum(d[1] for d in data) or 1
print(f"kernel {kern}: {tot} warp instructions, {tots} samples")
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print("%5.1f%% inst %5.1f%% samples  thr/inst %4.1f  %s:%-4s %s" % (100 * d[0] / tot, 100 * d[1] / tots, d[2] / max(d[0], 1), d[3], d[4], d[5].strip()[:110]))
