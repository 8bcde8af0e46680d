"""Policy / placement registries -- the reference's de-facto operator API
(/root/reference/core/scheduling/algorithm.py:182-187,292-298,442-444), backed
by the GPU: `placement_algorithms['yarn'](infrastructure, job, scheme)` scores
the job against the live node table with gs_place_batch and, like the
reference, COMMITS the reservation into the infrastructure on success."""
from __future__ import annotations

import numpy as np

from . import capi

_engine = None


def _eng(device=0):
    global _engine
    if _engine is None:
        _engine = capi.Engine(device=device, nsims=1)
    return _engine


def _lowest_bits(idle, cnt):
    take = 0
    for d in range(64):
        if cnt == 0:
            break
        if (idle >> d) & 1:
            take |= 1 << d
            cnt -= 1
    return take


def ms_yarn_placement(infrastructure, next_job, scheme):
    """-> (nodes: dict node_id -> NodeView, success)   (algorithm.py:28-32)"""
    cl = infrastructure.gs_cluster()
    tasks = int(next_job.task_count)
    gpc = int(next_job.gpu_per_worker)
    req = np.zeros(1, dtype=capi.JOBREQ_DTYPE)
    req[0] = (int(next_job.gpus), gpc, int(round(next_job.gpu_mem_max * 1048576)))
    first, used, task_node, _ = _eng().place_batch(cl, infrastructure.table, req, task_off=[0, tasks])
    if first[0] < 0:
        return {}, False
    nodes = {}
    tab = infrastructure.table
    gmask = (1 << infrastructure.num_gpu_p_node) - 1
    for t in range(tasks):
        nd = int(task_node[t])
        idle = ~int(tab["busy_mask"][nd]) & gmask
        tab["busy_mask"][nd] = int(tab["busy_mask"][nd]) | _lowest_bits(idle, gpc)
        tab["cpu_used"][nd] += cl.cpu_per_task
        tab["mem_used"][nd] += cl.mem_per_task
        next_job.tasks_running_on["%s_worker%d" % (next_job.job_id, t)] = str(nd + 1)
        nodes[str(nd + 1)] = infrastructure.nodes[str(nd + 1)]
    return nodes, True


def schedule_fifo(scheme, placement_algo, infrastructure, jobs_manager, delta, **kwargs):
    """(algorithm.py:189-202) kept for API shape; the engine runs the fused loop."""
    raise NotImplementedError("per-call fifo stepping is fused into Scheduler.start() on the device")


placement_algorithms = {"yarn": ms_yarn_placement}
scheduling_algorithms = {"fifo": schedule_fifo}
plugin_algorithms = {}
