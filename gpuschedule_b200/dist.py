"""Multi-GPU plumbing: one process per GPU, replicas sharded across ranks.

Two axes.  REPLICAS (throughput): rank r simulates its own slice of the sweep and nothing crosses
NVLink on the data path; torch.distributed (nccl on GPUs, gloo in the CPU tests) carries only the
start/stop barrier and the reduction of timers / counters to rank 0.  ONE SIMULATION ON SEVERAL GPUS
(BASELINE config C4, gittins): the per-event index evaluation is split over the ranks by chunks of
the runnable list (`chunk_owner`), results travel as peer stores inside the persistent kernel
(include/gsched.h: gs_comm_*); torch.distributed only distributes the 64-byte IPC handles
(`exchange_comm_handles`) before the run.
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


def chunk_owner(chunk: int, world: int) -> int:
    """Which rank evaluates chunk `chunk` (32 consecutive entries of the runnable list) of a sharded gittins event:
    round robin, the rule gs_sortpol_warp_kernel applies (gs_policy.cuh: `(base >> 5) % nr == me`)."""
    if world < 1 or chunk < 0:
        raise ValueError("bad chunk / world")
    return chunk % world


def exchange_comm_handles(handle: bytes, world: int, device=None):
    """All-gather of the ranks' 64-byte exchange-buffer handles (Engine.comm_prepare) -> list in rank order.
    Works with any torch.distributed backend: device tensors for nccl, cpu tensors for gloo."""
    if len(handle) != 64:
        raise ValueError("an IPC handle is 64 bytes")
    if world == 1:
        return [bytes(handle)]
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device or "cpu")
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]
