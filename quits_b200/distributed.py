"""Multi-GPU plumbing: one process per GPU, shots partitioned by global shot index, no data-path collective.

The only exchange is the end-of-run reduction of the per-rank error counters (sum) and device times (max) over
``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).  Because the noise of shot s depends only on
(seed, s), the union of the ranks' results is identical for every world size.
"""
from __future__ import annotations

import os

import numpy as np

from .engine import shard_range


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def reduce_results(counts: np.ndarray, times_ms, device=None):
    """Sum the counters and take the max of the timings over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    counts = np.asarray(counts, dtype=np.int64)
    times = np.asarray(times_ms, dtype=np.float64).reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts.copy(), times.copy()
    c = torch.from_numpy(counts.copy())
    t = torch.from_numpy(times.copy())
    if device is not None:
        c, t = c.to(device), t.to(device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return c.cpu().numpy(), t.cpu().numpy()


def my_shots(total: int):
    rank, ws, _ = world()
    return shard_range(int(total), rank, ws)
