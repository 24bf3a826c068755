"""Multi-GPU orchestration (one process per GPU, torch.distributed for the plumbing) — see slabs.py."""
