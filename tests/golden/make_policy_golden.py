#!/usr/bin/env python
"""Pin the event-driven policies against the reference's OWN (dead) loop code.

    python tests/golden/make_policy_golden.py        (build container only: needs /root/reference)

The reference keeps sjf / dlas / gittins only as functions that read module globals nobody defines
(JOBS, CLUSTER, LOG, scheduler -- SURVEY section 0), so they cannot run as shipped.  This script
executes those functions UNMODIFIED -- their source is taken from /root/reference/run_sim.py with
`ast` and exec'ed -- inside a namespace where the missing globals are small stubs that implement
exactly the completion documented in oracle/policy_oracle.c:
    JOBS      job_events (one start event per submit tick), runnable_jobs, queues, queue_limit,
              num_queue, gittins_delta, job_dist_data, move_to_runnable()
    CLUSTER   num_gpu / free_gpu, empty_infra(), release_job_res()  (infra/cluster.py:88-89)
    scheduler try_get_job_res() = the LIVE yarn placement rules on the emptied cluster
    LOG       job_complete(job, t), checkpoint(..., t): record what they are given
    FLAGS     schedule name;  util = the reference's own core/util.py
Everything the loops decide -- event selection, counter aging, MLFQ demotion, stable ordering,
greedy re-admission, preempt / resume marks, next completion, queue jumps, the gittins table and
index look-up -- is therefore the reference's code, not a restatement.  The recorded completions and
checkpoints are stored as fixtures (tests/golden/policy_*/) and compared with oracle/policy_oracle.c
by tests/test_policy_golden.py.
"""
import ast
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from gpuschedule_b200 import ingest, tracegen  # noqa: E402

WANTED = ["smallest_first_sim_jobs", "dlas_sim_jobs", "gittins_sim_jobs", "get_gittins_index",
          "cal_r_gittins_index", "parse_job_dist"]


def reference_functions(namespace):
    """exec the named function definitions of the reference's run_sim.py, verbatim, into `namespace`."""
    src = open(os.path.join(REF, "run_sim.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, "run_sim.py"), "exec"), namespace)
    return namespace


class Jobs:
    def __init__(self, jobs, num_queue=1, queue_limit=(), gittins_delta=3250):
        self.job_events = []
        for j in jobs:                                  # one start event per distinct submit tick, trace order
            if self.job_events and self.job_events[-1]["time"] == j["submit_time"]:
                self.job_events[-1]["start_jobs"].append(j)
            else:
                self.job_events.append({"time": j["submit_time"], "start_jobs": [j]})
        self.runnable_jobs = []
        self.num_queue = num_queue
        self.queues = [list() for _ in range(num_queue)]
        self.queue_limit = list(queue_limit)
        self.gittins_delta = gittins_delta
        self.job_dist_data = None

    def move_to_runnable(self, job):
        job["status"] = "PENDING"
        job["start_time"] = sys.maxsize
        job["last_check_time"] = job["submit_time"]
        job["total_executed_time"] = 0
        job["executed_time"] = 0
        job["pending_time"] = 0
        job["last_pending_time"] = 0
        self.runnable_jobs.append(job)


class Cluster:
    def __init__(self, M, G, K, fit_limit):
        self.M, self.G, self.K, self.fit_limit = M, G, K, fit_limit
        self.num_gpu = M * G
        self.empty_infra()

    def empty_infra(self):                              # infra/cluster.py:88-89 (+ node tables for sjf)
        self.free_gpu = self.num_gpu
        self.idle = [self.G] * self.M
        self.kfree = [self.K] * self.M

    def release_job_res(self, job):                     # everything is re-placed on every event anyway
        pass


class Scheduler:
    """try_get_job_res: the live yarn rules (core/scheduling/algorithm.py:28-32,301-417) on (idle, slots)."""

    @staticmethod
    def try_get_job_res(cluster, jobs, job):
        gpus, gpc = job["num_gpu"], job["gpc"]
        tasks = gpus // gpc
        if not (job["mem_bytes"] < cluster.fit_limit):
            return False
        if gpus <= cluster.G:
            for nd in range(cluster.M):
                if cluster.idle[nd] >= gpus and cluster.kfree[nd] >= tasks:
                    cluster.idle[nd] -= gpus
                    cluster.kfree[nd] -= tasks
                    return True
            return False
        caps = [min(cluster.idle[nd] // gpc, cluster.kfree[nd]) for nd in range(cluster.M)]
        if sum(c for c in caps if c > 0) < tasks:
            return False
        rem = tasks
        for nd in range(cluster.M):
            if caps[nd] <= 0 or rem == 0:
                continue
            take = min(caps[nd], rem)
            cluster.idle[nd] -= take * gpc
            cluster.kfree[nd] -= take
            rem -= take
        return True


class Log:
    def __init__(self, jobs_obj, cluster, counting):
        self.jobs_obj, self.cluster, self.counting = jobs_obj, cluster, counting
        self.completions, self.checkpoints = [], []

    def job_complete(self, job, t):
        self.completions.append([job["job_idx"], int(t), int(job["start_time"]), int(job["resume"]), int(job["preempt"])])

    def checkpoint(self, *args):
        t = int(args[-1])
        run = sum(1 for j in self.jobs_obj.runnable_jobs if j["status"] == "RUNNING")
        pend = sum(1 for j in self.jobs_obj.runnable_jobs if j["status"] == "PENDING")
        busy = sum(j["num_gpu"] for j in self.jobs_obj.runnable_jobs if j["status"] == "RUNNING")
        psum = sum(j["pending_time"] for j in self.jobs_obj.runnable_jobs if j["status"] == "PENDING")
        self.checkpoints.append([t, run, pend, busy, int(psum)])


def run_reference_policy(table, cluster_kw, policy, num_queue=1, queue_limit=(), gittins_delta=3250):
    M = cluster_kw["num_switch"] * cluster_kw["num_node_p_switch"]
    G = cluster_kw.get("num_gpu_p_node", 8)
    K = min(cluster_kw.get("num_cpu_p_node", 128) // 12, cluster_kw.get("mem_p_node", 512) // 60)
    fit_limit = (cluster_kw.get("gpu_memory_capacity", 32) * 1024 - 500) << 20
    need = np.maximum(1, np.ceil(table.duration)).astype(np.int64)
    jobs = [dict(job_idx=j, num_gpu=int(table.gpus[j]), gpc=int(table.gpu_per_task[j]), mem_bytes=int(table.mem_bytes[j]),
                 submit_time=int(table.arrive_tick[j]), duration=int(need[j]), status="ADDED", start_time=sys.maxsize,
                 end_time=0, preempt=0, resume=0, promote=0, q_id=0, rank=0, total_executed_time=0, executed_time=0,
                 pending_time=0, last_pending_time=0, last_check_time=0, remaining_time=0, remaining_gputime=0)
            for j in range(table.n)]
    JOBS = Jobs(jobs, num_queue, queue_limit, gittins_delta)
    CLUSTER = Cluster(M, G, K, fit_limit)
    LOG = Log(JOBS, CLUSTER, policy != "sjf")
    sys.path.insert(0, REF)
    from core import util as ref_util                     # the reference's own helpers (search_dict_list, print_fn)
    import copy
    import csv
    import math
    ns = dict(sys=sys, math=math, copy=copy, os=os, csv=csv, util=ref_util, JOBS=JOBS, CLUSTER=CLUSTER, LOG=LOG,
              scheduler=Scheduler, FLAGS=types.SimpleNamespace(schedule=policy), print=lambda *a, **k: None)
    reference_functions(ns)
    table_out = None
    if policy == "gittins":
        # parse_job_dist reads ./yarn-gput1000.csv with a 'duration' column (run_sim.py:1683-1708)
        d = tempfile.mkdtemp(prefix="gsgit_")
        with open(os.path.join(d, "yarn-gput1000.csv"), "w") as f:
            f.write("duration\n" + "\n".join(str(int(x)) for x in (need * table.gpus.astype(np.int64))) + "\n")
        cwd = os.getcwd()
        os.chdir(d)
        try:
            JOBS.job_dist_data = ns["parse_job_dist"]()
        finally:
            os.chdir(cwd)
        table_out = {"data": [float(x) for x in JOBS.job_dist_data["data"]],
                     "gittins": [float(x) for x in JOBS.job_dist_data["gittins"]]}
        ns["gittins_sim_jobs"](JOBS.job_dist_data, True, True)
    elif policy == "sjf":
        ns["smallest_first_sim_jobs"](False)
    else:
        ns["dlas_sim_jobs"](policy == "dlas-gpu", 0)
    unfinished = [[j["job_idx"], int(j["resume"]), int(j["preempt"])] for j in JOBS.runnable_jobs]
    return LOG.completions, LOG.checkpoints, table_out, unfinished


CASES = {
    "policy_sjf_sat": ("sjf", dict(num_switch=1, num_node_p_switch=6), dict(n=220, seed=71, rate=0.9), {}),
    "policy_sjf_wide": ("sjf", dict(num_switch=2, num_node_p_switch=8), dict(n=200, seed=72, rate=1.5, wide=True), {}),
    "policy_dlas_gpu": ("dlas-gpu", dict(num_switch=1, num_node_p_switch=6), dict(n=240, seed=73, rate=0.8),
                        dict(num_queue=4, queue_limit=[60, 200, 800])),
    "policy_dlas": ("dlas", dict(num_switch=1, num_node_p_switch=4), dict(n=200, seed=74, rate=0.8),
                    dict(num_queue=3, queue_limit=[20, 90])),
    "policy_gittins": ("gittins", dict(num_switch=1, num_node_p_switch=6), dict(n=220, seed=75, rate=0.9),
                       dict(gittins_delta=3250)),
    "policy_gittins_d200": ("gittins", dict(num_switch=1, num_node_p_switch=3), dict(n=260, seed=76, rate=1.2),
                            dict(gittins_delta=200)),
    "policy_dlas_gpu_8q": ("dlas-gpu", dict(num_switch=2, num_node_p_switch=3, num_gpu_p_node=4), dict(n=300, seed=77, rate=1.0),
                           dict(num_queue=8, queue_limit=[10, 30, 60, 120, 250, 500, 1000])),
    "policy_sjf_gpc2": ("sjf", dict(num_switch=1, num_node_p_switch=5, num_cpu_p_node=60, mem_p_node=300),
                        dict(n=240, seed=78, rate=1.0, gpc=2), {}),
    "policy_dlas_light": ("dlas", dict(num_switch=4, num_node_p_switch=8), dict(n=300, seed=79, rate=0.5),
                          dict(num_queue=2, queue_limit=[25])),
}


def build_table(n, seed, rate, wide=False, gpc=1):
    choices, probs = ([1, 2, 4, 8, 16, 24], [.3, .25, .2, .15, .06, .04]) if wide else ([1, 2, 4, 8], [.4, .3, .2, .1])
    if gpc > 1:
        choices = [c * gpc for c in choices]
    df = tracegen.synth_frame(n, seed=seed, rate=rate, gpu_choices=choices, gpu_probs=probs,
                              gpu_per_container=gpc, max_mem_mib=33000).drop(columns=["model"])
    return df


def main():
    for name, (policy, ckw, tkw, pkw) in CASES.items():
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        df = build_table(**tkw)
        trace = os.path.join(d, "trace.csv")
        df.to_csv(trace, index=False)
        table = ingest.JobTraceReader(trace).prepare_jobs().table(0.5)
        comp, chk, gtab, unfinished = run_reference_policy(table, ckw, policy, **pkw)
        meta = {"policy": policy, "cluster": ckw, "params": pkw,
                "reference": "run_sim.py functions exec'ed verbatim under tests/golden/make_policy_golden.py stubs"}
        json.dump(meta, open(os.path.join(d, "params.json"), "w"), indent=1, sort_keys=True)
        json.dump({"completions": comp, "checkpoints": chk, "gittins_table": gtab, "unfinished": unfinished}, open(os.path.join(d, "expected.json"), "w"))
        pre = sum(c[4] for c in comp)
        print(f"{name}: {len(comp)} completions, {len(chk)} checkpoints, {pre} preemptions")


if __name__ == "__main__":
    main()
