# dev: step traces with blocking and spinning in-step syncs (box noise check)
python tools/step_trace.py 10000000 60 > gpurun_out/trace_block.txt 2>&1
VGS_B200_SPIN_SYNC=1 python tools/step_trace.py 10000000 60 > gpurun_out/trace_spin.txt 2>&1
python tools/step_trace.py 10000000 60 > gpurun_out/trace_block2.txt 2>&1
python - <<'PY'
import ast, statistics
for f in ("gpurun_out/trace_block.txt", "gpurun_out/trace_spin.txt", "gpurun_out/trace_block2.txt"):
    t = ast.literal_eval(open(f).read().strip().splitlines()[-1])
    st = [x[0] for x in t[5:]]; gr = [x[1] for x in t[5:]]
    print(f, "step median %.1f mean %.2f max %.1f p90 %.1f | graph median %.1f mean %.2f max %.1f" % (statistics.median(st), statistics.mean(st), max(st), sorted(st)[int(len(st)*0.9)], statistics.median(gr), statistics.mean(gr), max(gr)))
PY
nproc; cat /proc/loadavg
