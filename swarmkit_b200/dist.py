"""Rank plumbing for the N>1 benchmark arm (one process per GPU, launched by torchrun).

The placement path's own collectives (per batch two small NCCL all-gathers and one all-reduce of the merged
class member lists of a node-sharded engine group) and the split of the node axis over
the ranks live inside libplacement.so; what crosses
ranks HERE is the timing reduction the bench contract asks for: MAX over ranks of the device
time, SUM over ranks of the work done (the ranks of a group hold replicas of the same
decisions, so bench.py lets only rank 0 contribute work)."""
from __future__ import annotations

import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def reduce_step(times, counts, device=None):
    """times -> elementwise MAX over ranks, counts -> elementwise SUM over ranks (lists of floats)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(times), list(counts)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(times), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()
