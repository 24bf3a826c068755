"""Multi-GPU path.  not-gpu: the range partition and the connect-list exchange protocol over a
world_size-2 gloo group on CPU (fake handles carrying the oracle's lists); gpu (needs >= 2 GPUs):
the partitioned run must reproduce the single-GPU labels bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unit_ranges_cover_and_balance():
    from vgs_svgs_segmentation_b200.multigpu import unit_ranges
    rng = np.random.default_rng(0)
    n = rng.integers(1, 120, 5000)
    off = np.concatenate([[0], np.cumsum(n)])
    for world in (1, 2, 3, 8):
        r = unit_ranges(off, world)
        assert r[0][0] == 0 and r[-1][1] == 5000 and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        work = [float((n[a:b].astype(np.float64) ** 2).sum()) for a, b in r]
        assert max(work) <= 1.2 * sum(work) / world + 120.0 ** 2
    assert unit_ranges(np.array([0, 3]), 4)[-1][1] == 1     # more ranks than units


def _gloo_worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from vgs_svgs_segmentation_b200 import scenes
    from vgs_svgs_segmentation_b200.multigpu import exchange_connect, unit_ranges
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    xyz = scenes.two_planes(20_000, seed=3)
    r = oracle.run(xyz, math=1)
    adj_off = r.adj_offsets
    nu = len(adj_off) - 1
    ranges = unit_ranges(adj_off, world)
    slots = [(int(adj_off[a]), int(adj_off[b])) for a, b in ranges]
    # "device" state of this rank: only its own range of the connect lists is filled
    full_cnt = np.diff(r.conn0_offsets).astype(np.int32)
    full_idx = np.full(int(adj_off[-1]), -1, np.int32)
    for u in range(nu):
        c = full_cnt[u]
        full_idx[adj_off[u]:adj_off[u] + c] = r.conn0_idx[r.conn0_offsets[u]:r.conn0_offsets[u] + c]
    a, b = ranges[rank]
    cnt = np.zeros(nu, np.int32); idx = np.full(int(adj_off[-1]), -1, np.int32)
    cnt[a:b] = full_cnt[a:b]; idx[slots[rank][0]:slots[rank][1]] = full_idx[slots[rank][0]:slots[rank][1]]

    def export_fn(f, l, c, i):
        c[:l - f] = torch.from_numpy(cnt[f:l]); e0, e1 = int(adj_off[f]), int(adj_off[l]); i[:e1 - e0] = torch.from_numpy(idx[e0:e1])

    def import_fn(f, l, c, i):
        cnt[f:l] = c[:l - f].numpy(); e0, e1 = int(adj_off[f]), int(adj_off[l]); idx[e0:e1] = i[:e1 - e0].numpy()

    exchange_connect(ranges, slots, rank, export_fn, import_fn,
                     new_tensor=lambda n: torch.zeros(max(int(n), 1), dtype=torch.int32),
                     broadcast=lambda t, src: dist.broadcast(t, src=src))
    ok = np.array_equal(cnt, full_cnt) and np.array_equal(idx, full_idx)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_exchange_protocol_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 1000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


@pytest.mark.gpu
def test_partitioned_two_gpus_equals_single(built_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py"), "400000"],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout


@pytest.mark.gpu
def test_library_unit_ranges_equal_python_rule(built_lib):
    """vgs_unit_ranges (evaluated inside the library, no numpy on the step path) = unit_ranges() on the same offsets"""
    from vgs_svgs_segmentation_b200 import capi, scenes
    from vgs_svgs_segmentation_b200.multigpu import unit_ranges
    xyz = scenes.construction_site(150_000, seed=4, extent=9.0)
    h = capi.Handle()
    h.set_points(xyz)
    h.voxelize(0.15)
    h.compute_features(10)
    h.find_adjacency(0.5)
    off = h.blob("ADJ_OFFSETS")
    for world in (1, 2, 3, 8):
        ranges, slots = h.unit_ranges(world)
        assert ranges == unit_ranges(off, world)
        assert slots == [(int(off[a]), int(off[b])) for a, b in ranges]
