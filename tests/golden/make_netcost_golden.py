#!/usr/bin/env python
"""Golden vectors for the network-cost model, made by the reference's own function (build container only).

    python tests/golden/make_netcost_golden.py

core/network/network_service.py imports fine; what the reference lacks is a Job with the attributes the
function reads (ps_count, model_size, iterations, PS task ids -- SURVEY appendix A.7).  This script calls
the UNMODIFIED calculate_network_costs(infrastructure, job) with a stub job / infrastructure carrying
exactly those attributes (is_distributed() restated from core/jobs/job.py:199-200) on 400 random cases
and stores inputs + the returned float (as float.hex, so the comparison is bit-exact) in
tests/golden/netcost.json.  tests/test_oracle_golden.py checks oracle_net_cost against it on CPU,
tests/test_gpu_parity.py checks gs_net_cost against it on the GPU.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class StubJob:
    def __init__(self, job_id, ps_count, model_size, iterations, tasks_running_on):
        self.job_id, self.ps_count, self.model_size, self.iterations = job_id, ps_count, model_size, iterations
        self.tasks_running_on = tasks_running_on

    def is_distributed(self):                           # core/jobs/job.py:199-200
        return self.ps_count > 1


def main():
    sys.path.insert(0, REF)
    from core import util as ref_util
    from core.network import network_service
    ref_util.print_fn = lambda *a, **k: None            # silence the per-call log line
    rng = np.random.default_rng(20260921)
    sizes_mb = [15.0, 97.49, 233.1, 548.0, 1300.0, 0.5, 5000.25]
    iters = [1, 109, 521, 4861, 28000, 2.5]
    cases = []
    for i in range(400):
        bandwidth = float(rng.choice([1250.0, 1250.0, 12500.0, 100.0, 3333.3]))
        latency = float(rng.choice([0.015, 0.015, 0.0005, 0.1, 0.0]))
        infra = types.SimpleNamespace(bandwidth=bandwidth, internode_latency=latency)
        n_tasks = int(rng.integers(1, 40))
        n_nodes = int(rng.integers(1, 9))
        node = rng.integers(0, n_nodes, size=n_tasks).tolist()
        ps_frac = float(rng.choice([0.0, 0.2, 0.5, 1.0]))
        is_ps = (rng.random(n_tasks) < ps_frac).astype(int).tolist()
        if rng.random() < 0.15:                          # PS co-located with every worker node: cost 0 path
            node = node + node
            is_ps = [0] * n_tasks + [1] * n_tasks
        ps_count = int(rng.integers(0, 5))
        model = float(rng.choice(sizes_mb))
        it = float(rng.choice(iters))
        tasks = {("ps_%d" % t if p else "worker_%d" % t): str(nd + 1) for t, (nd, p) in enumerate(zip(node, is_ps))}
        out = network_service.calculate_network_costs(infra, StubJob("j%d" % i, ps_count, model, it, tasks))
        cases.append(dict(bandwidth=bandwidth, latency=latency, node=node, is_ps=is_ps, ps_count=ps_count,
                          model_mb=model, iterations=it, expected=float(out).hex()))
    with open(os.path.join(HERE, "netcost.json"), "w") as f:
        json.dump({"reference": "core/network/network_service.py:3-39 called unmodified with stub job / infrastructure",
                   "cases": cases}, f)
    nz = sum(1 for c in cases if float.fromhex(c["expected"]) != 0.0)
    print(f"{len(cases)} cases, {nz} with a non-zero cost")


if __name__ == "__main__":
    main()
