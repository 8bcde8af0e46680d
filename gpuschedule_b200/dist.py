"""Multi-GPU plumbing: one process per GPU, replicas sharded across ranks.

A single simulation does not shard usefully (one shared cluster state, at most one placement
per simulated tick, every tick ~1 us of work on one GPU versus >= 10 us for any cross-GPU
exchange), so the multi-GPU axis is REPLICAS: rank r simulates its own slice of the sweep
and nothing crosses NVLink on the data path.  torch.distributed (nccl on GPUs, gloo in the
CPU tests) is used only for the start/stop barrier and for reducing the timing / event
counters to rank 0.
"""
from __future__ import annotations

import os


def env_rank():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def replica_seeds(rank: int, world: int, per_rank: int, base: int = 1):
    """Seeds of the traces rank `rank` simulates: a contiguous, disjoint slice per rank
    (weak scaling: per-rank work is fixed)."""
    if not (0 <= rank < world) or per_rank < 0:
        raise ValueError("bad rank / world / per_rank")
    lo = base + rank * per_rank
    return list(range(lo, lo + per_rank))


def shard_range(n_items: int, rank: int, world: int):
    """[lo, hi) of a strong-scaling split of `n_items` sweep points over `world` ranks."""
    q, r = divmod(n_items, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class Reducer:
    """max / sum of python floats over ranks (device tensors for nccl, cpu tensors for gloo)."""

    def __init__(self, world: int, device=None):
        self.world = world
        self.device = device

    def _reduce(self, x, op):
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device or "cpu")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.MAX)

    def sum(self, x):
        import torch.distributed as dist
        return self._reduce(x, dist.ReduceOp.SUM)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
