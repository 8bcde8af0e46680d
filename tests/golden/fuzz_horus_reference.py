#!/usr/bin/env python
"""Differential fuzz of oracle/horus_oracle.c against the UNMODIFIED reference (build container only).

    python tests/golden/fuzz_horus_reference.py [n_cases] [first_seed]

Every case draws a random cluster, trace, (scheme, schedule) in {horus, horus+, gandiva}, look-ahead
width, queue count and numpy seed, runs /root/reference/run_sim.py through make_golden.run_reference and
compares job.csv and all 13 columns of cluster.csv byte for byte with the restatement.  A reference run
that raises is reported and skipped.  Needs /root/reference, so it is NOT part of the suite; its last run
is recorded in DESIGN.md section 5.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from conftest import render_horus_outputs  # noqa: E402
from gpuschedule_b200 import capi, ingest, tracegen  # noqa: E402
import oracle  # noqa: E402


def random_case(seed):
    rng = np.random.default_rng(7000 + seed)
    kind = str(rng.choice(["horus", "horus", "gandiva", "horus+"]))
    sched = kind
    if seed >= 1000:                                   # cross combinations of score function and scheduler
        sched = str(rng.choice(["fifo", "horus", "gandiva", "horus+"]))
    if seed >= 2000:                                   # --scheme yarn (no packing) under these schedulers
        kind, sched = "yarn", str(rng.choice(["horus", "gandiva", "horus+"]))
    G = int(rng.choice([2, 4, 8, 8]))
    flags = dict(num_switch=int(rng.integers(1, 4)), num_node_p_switch=int(rng.integers(1, 6)), num_gpu_p_node=G,
                 num_cpu_p_node=int(rng.choice([36, 60, 128, 128])), mem_p_node=int(rng.choice([180, 300, 512, 512])),
                 gpu_memory_capacity=int(rng.choice([16, 32, 32])), _scheme=kind, _schedule=sched,
                 num_buffer=int(rng.choice([1, 2, 5, 15])))
    if sched == "horus+":
        flags["num_queue"] = int(rng.integers(2, 6))
    gpc = int(rng.choice([1, 1, 1, 2])) if G >= 2 else 1
    choices = sorted(set(int(x) * gpc for x in rng.choice([1, 1, 2, 2, 3, 4, 6, 8, 12], size=4)))
    probs = rng.dirichlet(np.ones(len(choices)))
    n = int(rng.integers(15, 90))
    df = tracegen.synth_frame(n, seed=3000 + seed, rate=float(rng.choice([0.5, 1.0, 2.0, 4.0])), gpu_per_container=gpc,
                              gpu_choices=choices, gpu_probs=probs,
                              max_mem_mib=int(rng.choice([6000, 12000, 16384, 33500]))).drop(columns=["model"])
    if rng.random() < 0.3:
        df["minutes"] = np.round(df["minutes"] * float(rng.choice([0.3, 3.0])), 3)     # short jobs / long jobs (time slices)
    if rng.random() < 0.2:
        df.loc[rng.choice(n, max(1, n // 10), replace=False), "gpu_utilization_avg"] = 0.0     # Job.__lt__ falsy branch
        df["gpu_utilization_max"] = np.maximum(df["gpu_utilization_max"], df["gpu_utilization_avg"])
    return df, flags, int(rng.integers(0, 2 ** 31 - 1))


def run_case(seed, workdir):
    df, flags, np_seed = random_case(seed)
    d = os.path.join(workdir, f"c{seed}")
    os.makedirs(d)
    trace = os.path.join(d, "trace.csv")
    df.to_csv(trace, index=False)
    old = mg.SEED
    mg.SEED = np_seed
    try:
        mg.run_reference(trace, flags, d)
    except Exception as e:                                              # noqa: BLE001
        return f"reference raised ({type(e).__name__})", flags
    finally:
        mg.SEED = old
    cl = {k: v for k, v in flags.items() if not k.startswith("_") and k not in ("num_buffer", "num_queue")}
    cluster = capi.make_cluster(**cl)
    table = ingest.JobTraceReader(trace).prepare_jobs().table(0.5)
    res = oracle.run_horus(cluster, table, scheme=flags["_scheme"], schedule=flags["_schedule"],
                           num_buffer=flags["num_buffer"], num_queue=flags.get("num_queue", 1), seed=np_seed)
    got_job, got_cluster = render_horus_outputs(table, cluster, res)
    exp_job = open(os.path.join(d, "job.csv"), newline="").read()
    exp_cluster = open(os.path.join(d, "cluster.csv"), newline="").read()
    shutil.rmtree(d)
    if got_job != exp_job or got_cluster != exp_cluster:
        gl, el = got_cluster.split("\r\n"), exp_cluster.split("\r\n")
        first = next((i for i, (a, b) in enumerate(zip(gl, el)) if a != b), min(len(gl), len(el)))
        return f"MISMATCH (job.csv equal: {got_job == exp_job}; first cluster.csv line {first} of {len(el)})", flags
    return None, flags


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ok = bad = raised = 0
    with tempfile.TemporaryDirectory(prefix="gshorusfuzz_") as wd:
        for seed in range(first, first + n_cases):
            msg, flags = run_case(seed, wd)
            if msg is None:
                ok += 1
            elif msg.startswith("reference raised"):
                raised += 1
                print(f"seed {seed}: {msg} {flags}", flush=True)
            else:
                bad += 1
                print(f"seed {seed}: {msg} {flags}", flush=True)
    print(f"{ok} cases identical, {bad} different, {raised} where the reference raised")


if __name__ == "__main__":
    main()
