"""bench.py contract pieces that can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads (impl, metric, value, unit, e2e, cpu_baseline ...), ranks > 0 stay silent, and the default arm refuses to
run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                          env={**os.environ, **(env or {})})


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-points", "20000"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "points/sec segmented end-to-end" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "20000 points" in cb["sample"]
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-points", "20000", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--points", "20000", "--no-cpu"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
