"""CPU: host side of the multi-GPU slab split — slab cuts, slices, and the N > 1 plumbing on a
world_size-2 gloo group (the NCCL unique id travels over torch.distributed; kernels need a GPU and are covered by
tests/test_slabs_gpu.py)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_bounds_cover_the_cloud(built_lib):
    from vgs_svgs_segmentation_b200 import slabs
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            b = [slabs.slice_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_choose_cuts_balances_points_on_the_widest_axis(built_lib):
    from vgs_svgs_segmentation_b200 import capi
    h = np.zeros((3, 256), np.uint64)
    h[0, 10:210] = 100          # x: 200 occupied bins
    h[1, 50:90] = 500           # y: 40 bins
    h[2, 3] = 20000             # z: a flat ground
    axis, cuts = capi.choose_cuts(h, 0, 4)
    assert axis == 0
    assert cuts[0] < 0 and cuts[-1] > 255
    assert list(cuts[1:-1]) == [60, 110, 160]
    # bins of 4 keys (shift 2): cuts are multiples of 4
    axis, cuts = capi.choose_cuts(h, 2, 2)
    assert axis == 0 and cuts[1] == 110 * 4


def test_choose_cuts_degenerate_inputs(built_lib):
    from vgs_svgs_segmentation_b200 import capi
    h = np.zeros((3, 64), np.uint64)
    h[1, 7] = 5                 # every point in one bin: all but one slab stay empty, cuts never decrease
    axis, cuts = capi.choose_cuts(h, 0, 8)
    assert all(cuts[i] <= cuts[i + 1] for i in range(8))
    # one rank: no interior cut
    axis, cuts = capi.choose_cuts(h, 0, 1)
    assert len(cuts) == 2 and cuts[0] < 0 < cuts[1]
    # uneven histogram: every slab gets its share to within one bin
    rng = np.random.default_rng(3)
    h = np.zeros((3, 512), np.uint64)
    h[2] = rng.integers(0, 1000, 512)
    axis, cuts = capi.choose_cuts(h, 0, 8)
    assert axis == 2
    tot = int(h[2].sum())
    edges = [0] + [int(c) for c in cuts[1:-1]] + [512]
    for r in range(8):
        share = int(h[2, edges[r]:edges[r + 1]].sum())
        assert abs(share - tot / 8) <= 1000 + 1


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    from vgs_svgs_segmentation_b200 import slabs
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    uid = slabs.share_unique_id(dist, rank)
    import torch
    t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).clone()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert bytes(t.numpy().tobytes()) == bytes(ref.numpy().tobytes()) and len(uid) == 128 and any(uid)
    first, last = slabs.slice_bounds(1001, world, rank)
    tot = torch.tensor([last - first])
    dist.all_reduce(tot)
    assert int(tot) == 1001
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_unique_id_travels_over_gloo_world_size_2(built_lib, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
