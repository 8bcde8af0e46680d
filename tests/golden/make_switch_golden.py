#!/usr/bin/env python
"""Pin the legacy switch-local yarn placement + PS/worker traffic accounting against the reference's OWN code.

    python tests/golden/make_switch_golden.py        (build container only: needs /root/reference)

SURVEY row a13, second half: `_Cluster.ms_yarn_placement` (infra/cluster.py:888-898) tries the switches in order
and `_Switch.ms_yarn_alloc_res -> try_cross_node_alloc / try_single_node_alloc` (infra/switch.py:38-167,190-206)
place a job inside ONE switch, charging cpus / memory and the per-node network load of the parameter-server
shards.  infra/switch.py cannot be imported (it needs the non-existent `core.job` and a `_Node` class nobody
defines), so -- exactly like make_policy_golden.py does for the policy loops -- the METHOD SOURCES are taken out
of the reference with `ast` and executed UNMODIFIED; what the repository never defines is supplied by stubs:

    _Node      id, num_gpu, free_gpus, free_cpus, free_mem, network_in/out; check_free_gpus(), check_free_cpus(),
               alloc_job_res(num_gpu, num_cpu) -> False and no change if either is short, else charge both;
               add_network_load(in, out) adds to the two counters
    job_queue  worker_mem / ps_mem / p_w_mem = 5 / 8 / 0.2 (core/models.py:24-26, the only live definition;
               job_queue_manager.py:23-25 has a commented-out second set), create_multi_nodes_placement(job,
               switch_id, node_list) and create_single_node_placement(job, switch_id, node_id, gpu, cpu, mem)
               record their arguments (the reference defines neither)
    job        {'num_gpu', 'model': {'total_size'}, 'ps_network': [...]}: ps_network (one traffic figure per
               parameter-server shard, produced nowhere in the repository) is an INPUT of the fixtures

Every decision and every number -- which switch, which nodes, cpus and memory per node, the traffic expression
with its round(., 1) after every shard (switch.py:98-108,122-133) -- is therefore the reference's own code.
Fixtures: tests/golden/switch_yarn.json (cluster states, jobs, the reference's answers and the node tables after).
"""
import ast
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def reference_methods():
    """{name: function} of _Switch.try_cross_node_alloc / try_single_node_alloc / ms_yarn_alloc_res and
    _Cluster.ms_yarn_placement, compiled verbatim from the reference sources."""
    out = {}
    for path, cls, names in ((os.path.join(REF, "infra", "switch.py"), "_Switch",
                              ["try_cross_node_alloc", "try_single_node_alloc", "ms_yarn_alloc_res"]),
                             (os.path.join(REF, "infra", "cluster.py"), "_Cluster", ["ms_yarn_placement"])):
        tree = ast.parse(open(path).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == cls:
                for fn in node.body:
                    if isinstance(fn, ast.FunctionDef) and fn.name in names:
                        ns = {"math": math}
                        exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
                        out[fn.name] = ns[fn.name]
    assert len(out) == 4, sorted(out)
    return out


class Node:
    def __init__(self, id, num_gpu, free_gpus, free_cpus, free_mem):
        self.id, self.num_gpu = id, num_gpu
        self.free_gpus, self.free_cpus, self.free_mem = free_gpus, free_cpus, free_mem
        self.network_in = self.network_out = 0.0

    def check_free_gpus(self):
        return self.free_gpus

    def check_free_cpus(self):
        return self.free_cpus

    def alloc_job_res(self, num_gpu, num_cpu):
        if num_gpu > self.free_gpus or num_cpu > self.free_cpus:
            return False
        self.free_gpus -= num_gpu
        self.free_cpus -= num_cpu
        return True

    def add_network_load(self, in_load=0, out_load=0):
        self.network_in += in_load
        self.network_out += out_load


class JobQueue:
    worker_mem, ps_mem, p_w_mem = 5, 8, 0.2            # core/models.py:24-26

    def __init__(self):
        self.placements = []

    def create_multi_nodes_placement(self, job, switch_id, node_list):
        self.placements.append({"switch": switch_id, "nodes": [{k: v for k, v in nd.items() if k != "tasks"} for nd in node_list]})

    def create_single_node_placement(self, job, switch_id, node_id, num_gpu, num_cpu, mem):
        self.placements.append({"switch": switch_id, "nodes": [{"id": node_id, "num_gpu": num_gpu, "num_cpu": num_cpu, "mem": mem,
                                                                  "network": None}]})
        return 0


class Switch:
    def __init__(self, id, nodes, num_gpu_p_node, methods):
        self.id, self.node_list, self.num_gpu_p_node = id, nodes, num_gpu_p_node
        self._m = methods

    def try_cross_node_alloc(self, job_queue, job):
        return self._m["try_cross_node_alloc"](self, job_queue, job)

    def try_single_node_alloc(self, job_queue, job):
        return self._m["try_single_node_alloc"](self, job_queue, job)

    def ms_yarn_alloc_res(self, job_queue, job):
        return self._m["ms_yarn_alloc_res"](self, job_queue, job)


class Cluster:
    def __init__(self, switches, methods):
        self.switch_list = switches
        self._m = methods

    def ms_yarn_placement(self, job_queue, job):
        return self._m["ms_yarn_placement"](self, job_queue, job)


def make_case(rng, idx):
    S = int(rng.integers(1, 5))
    P = int(rng.integers(2, 9))
    G = int(rng.choice([4, 8]))
    cpus, mem = int(rng.choice([64, 96, 128])), float(rng.choice([128, 256, 512]))
    load = float(rng.uniform(0.0, 0.8))
    nodes = []
    for s in range(S):
        for p in range(P):
            used = int(rng.integers(0, G + 1)) if rng.random() < load else 0
            nodes.append({"free_gpus": G - used, "free_cpus": cpus - int(rng.integers(0, 7)) * used,
                          "free_mem": mem - float(np.round(rng.uniform(0, 12), 1)) * used})
    jobs = []
    for _ in range(int(rng.integers(4, 12))):
        g = int(rng.choice([1, 1, 2, 4, 8, 8, 12, 16, 20, 24, 32]))
        total = float(np.round(rng.uniform(5, 600), 1))
        if g == 1 and rng.random() < 0.6:
            ps = []
        else:
            cut = np.sort(rng.uniform(0, total, size=g - 1)) if g > 1 else np.zeros(0)
            ps = [float(np.round(x, 1)) for x in np.diff(np.concatenate([[0.0], cut, [total]]))]
        jobs.append({"num_gpu": g, "total_size": total, "ps_network": ps})
    return {"name": f"case{idx}", "num_switch": S, "num_node_p_switch": P, "num_gpu_p_node": G, "nodes": nodes, "jobs": jobs}


def run_case(case, methods):
    S, P, G = case["num_switch"], case["num_node_p_switch"], case["num_gpu_p_node"]
    switches = []
    for s in range(S):
        nl = [Node(p, G, **case["nodes"][s * P + p]) for p in range(P)]
        switches.append(Switch(s, nl, G, methods))
    cluster, jq = Cluster(switches, methods), JobQueue()
    answers = []
    for jd in case["jobs"]:
        job = {"num_gpu": jd["num_gpu"], "model": {"total_size": jd["total_size"]}, "ps_network": list(jd["ps_network"])}
        before = len(jq.placements)
        ok = cluster.ms_yarn_placement(jq, job)
        ans = {"ok": bool(ok)}
        if ok:
            assert len(jq.placements) == before + 1
            pl = jq.placements[-1]
            ans["switch"] = pl["switch"]
            ans["nodes"] = [{"id": nd["id"], "num_gpu": nd["num_gpu"], "num_cpu": nd["num_cpu"], "mem": float(nd["mem"]).hex(),
                             "network": None if nd["network"] is None else float(nd["network"]).hex()} for nd in pl["nodes"]]
        answers.append(ans)
    after = [{"free_gpus": nd.free_gpus, "free_cpus": nd.free_cpus, "free_mem": float(nd.free_mem).hex(),
              "network_in": float(nd.network_in).hex()} for sw in switches for nd in sw.node_list]
    return answers, after


def main():
    methods = reference_methods()
    rng = np.random.default_rng(20260921)
    cases = []
    for i in range(60):
        case = make_case(rng, i)
        case["answers"], case["after"] = run_case(case, methods)
        cases.append(case)
    placed = sum(a["ok"] for c in cases for a in c["answers"])
    total = sum(len(c["answers"]) for c in cases)
    cross = sum(1 for c in cases for a in c["answers"] if a["ok"] and len(a["nodes"]) > 1)
    out = {"reference": "matthewygf/GPUSchedule @ ea0f1474: infra/switch.py:38-167,190-206 and infra/cluster.py:888-898 executed verbatim under the stubs documented in make_switch_golden.py",
           "mem_constants": {"worker_mem": 5, "ps_mem": 8, "p_w_mem": 0.2}, "cases": cases}
    with open(os.path.join(HERE, "switch_yarn.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(f"{len(cases)} cases, {total} jobs, {placed} placed, {cross} across several nodes")


if __name__ == "__main__":
    main()
