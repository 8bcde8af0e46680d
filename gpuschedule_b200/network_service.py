"""calculate_network_costs -- drop-in for
/root/reference/core/network/network_service.py:3-39, evaluated on the GPU
(gs_net_cost) for one job or a batch."""
from __future__ import annotations

import numpy as np

from . import capi

_engine = None


def _eng(device=0):
    global _engine
    if _engine is None:
        _engine = capi.Engine(device=device, nsims=1)
    return _engine


def calculate_network_costs(infrastructure, job):
    if not job.is_distributed():
        return 0
    keys = list(job.tasks_running_on.keys())
    node = np.array([int(job.tasks_running_on[k]) - 1 for k in keys], dtype=np.int32)
    is_ps = np.array([1 if "ps" in k else 0 for k in keys], dtype=np.uint8)
    out = _eng().net_cost(infrastructure.gs_cluster(), [0, len(keys)], node, is_ps, [job.ps_count],
                          [job.model_size], [job.iterations])
    return float(out[0]) if out[0] != 0.0 else 0
