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
    held = getattr(next_job, "held", None)
    for t in range(tasks):
        nd = int(task_node[t])
        idle = ~int(tab["busy_mask"][nd]) & gmask
        take = _lowest_bits(idle, gpc)
        tab["busy_mask"][nd] = int(tab["busy_mask"][nd]) | take
        tab["cpu_used"][nd] += cl.cpu_per_task
        tab["mem_used"][nd] += cl.mem_per_task
        if held is not None:
            held.append((nd, take))
        next_job.tasks_running_on["%s_worker%d" % (next_job.job_id, t)] = str(nd + 1)
        nodes[str(nd + 1)] = infrastructure.nodes[str(nd + 1)]
    return nodes, True


def schedule_fifo(scheme, placement_algo, infrastructure, jobs_manager, delta, **kwargs):
    """-> (nodes or None, job or None, success or None)   (algorithm.py:189-202): one attempt on the queue head; the
    placement itself is the GPU call behind `placement_algo`.  Scheduler.start() runs the fused loop on the device;
    this entry point is the same step driven from the host, one call per tick, like the reference drives it."""
    next_job = jobs_manager.get_next_job(delta)
    if next_job is None:
        return None, None, None
    nodes, success = placement_algo(infrastructure, next_job, scheme)
    if success:
        jobs_manager.pop(delta)
    return nodes, next_job, success


def release_job(infrastructure, job):
    """Node.release_allocated_resources for every task of a finished job (node.py:71-91): devices, cpu and memory of
    what ms_yarn_placement committed go back to the node table."""
    cl = infrastructure.gs_cluster()
    tab = infrastructure.table
    for nd, take in job.held:
        tab["busy_mask"][nd] = int(tab["busy_mask"][nd]) & ~take
        tab["cpu_used"][nd] -= cl.cpu_per_task
        tab["mem_used"][nd] -= cl.mem_per_task
    job.held = []


placement_algorithms = {"yarn": ms_yarn_placement}
scheduling_algorithms = {"fifo": schedule_fifo}
# the reference registers only gandiva's time-slice hook here (algorithm.py:442-444); that policy runs inside the
# utilisation-aware engine (schedule.py: _start_utilisation_aware), so nothing is hooked per tick on this side
plugin_algorithms = {}
