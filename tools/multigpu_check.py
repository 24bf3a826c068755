"""torchrun --nproc-per-node N tools/multigpu_check.py [points]: the partitioned N-GPU run must give the
single-GPU labels bit for bit (checked on every rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as g
g.build(oracle=False, quiet=True)
from vgs_svgs_segmentation_b200 import capi, scenes
from vgs_svgs_segmentation_b200.multigpu import segment_partitioned

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
xyz = scenes.construction_site(n, seed=3, extent=70.0 * (n / 10_000_000) ** 0.5)
params = capi.make_params()
h = capi.Handle(device=local, stream=torch.cuda.current_stream().cuda_stream)
h.set_points(xyz)
single = h.run(params).copy()
h.set_points(xyz)
multi = segment_partitioned(h, params, rank, world)
same = bool(np.array_equal(single, multi))
t = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} points={n} clusters={h.counts()['n_clusters_exported']} labels_equal_single_gpu={bool(t.item())}")
    print("MULTIGPU_OK" if t.item() == 1 else "MULTIGPU_MISMATCH")
h.close()
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
