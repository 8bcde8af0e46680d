#!/usr/bin/env python
"""Differential fuzz of the CPU oracle against the UNMODIFIED reference (build container only).

    python tests/golden/fuzz_reference.py [n_cases] [first_seed]

Every case draws a random cluster (switches x nodes x GPUs, cpu / mem per node, GPU memory) and a
random small trace (arrival rate, GPU mix, gpu_per_container, occasional over-sized memory, ties,
fractional times, filtered / NaN rows), runs /root/reference/run_sim.py on it through the same
harness as make_golden.py and compares job.csv and all 13 columns of cluster.csv byte for byte with
oracle/gsched_oracle.c + the package's formatter.  It needs /root/reference, so it is NOT a test
of the suite; its last run is recorded in DESIGN.md section 5.
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

import make_golden  # noqa: E402
from conftest import render_outputs  # noqa: E402
from gpuschedule_b200 import capi, ingest, tracegen  # noqa: E402
import oracle  # noqa: E402


def random_case(seed):
    rng = np.random.default_rng(seed)
    G = int(rng.choice([2, 4, 8, 8, 8, 16]))
    flags = dict(num_switch=int(rng.integers(1, 4)), num_node_p_switch=int(rng.integers(1, 13)), num_gpu_p_node=G,
                 num_cpu_p_node=int(rng.choice([24, 60, 128, 128, 200])), mem_p_node=int(rng.choice([120, 300, 512, 512, 1000])),
                 gpu_memory_capacity=int(rng.choice([16, 32, 32])))
    gpc = int(rng.choice([1, 1, 1, 2])) if G >= 2 else 1
    choices = sorted(set(int(x) * gpc for x in rng.choice([1, 1, 2, 2, 3, 4, 6, 8, 12, 16, 24], size=5)))
    probs = rng.dirichlet(np.ones(len(choices)))
    n = int(rng.integers(20, 160))
    df = tracegen.synth_frame(n, seed=1000 + seed, rate=float(rng.choice([0.3, 0.8, 1.5, 3.0])), gpu_per_container=gpc,
                              gpu_choices=choices, gpu_probs=probs,
                              max_mem_mib=int(rng.choice([8000, 16384, 16384, 17000, 33500]))).drop(columns=["model"])
    if rng.random() < 0.5:
        df["normalized_time"] = df["normalized_time"] + rng.integers(0, 4, size=n) * 2500 + int(rng.integers(0, 10 ** 6))
    if rng.random() < 0.4:
        df.loc[rng.choice(n, max(1, n // 15), replace=False), "type"] = "interactive"
    if rng.random() < 0.3:
        df["minutes"] = df["minutes"].astype(float)
        df.loc[rng.choice(n, max(1, n // 20), replace=False), "minutes"] = np.nan
    if rng.random() < 0.5:
        df = df.sample(frac=1.0, random_state=int(seed))
    return df, flags


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    bad = 0
    for seed in range(first, first + n_cases):
        df, flags = random_case(seed)
        d = tempfile.mkdtemp(prefix="gsfuzz_")
        trace = os.path.join(d, "trace.csv")
        df.to_csv(trace, index=False)
        try:
            make_golden.run_reference(trace, flags, d)
        except Exception as exc:                       # the reference itself crashed (outside its input domain)
            print(f"seed {seed}: reference failed ({type(exc).__name__}), skipped")
            shutil.rmtree(d)
            continue
        table = ingest.JobTraceReader(trace).prepare_jobs().table(0.5)
        cluster = capi.make_cluster(**flags)
        r = oracle.run_fifo(cluster, table)
        job_csv, cluster_csv = render_outputs(table, cluster, r.rows, r.recs, r.finish_order, r.span_off, r.spans,
                                              make_golden.SEED)
        exp_job = open(os.path.join(d, "job.csv"), newline="").read()
        exp_cluster = open(os.path.join(d, "cluster.csv"), newline="").read()
        ok = job_csv == exp_job and cluster_csv == exp_cluster
        print(f"seed {seed}: M={cluster.num_switch * cluster.num_node_p_switch} G={flags['num_gpu_p_node']} n={table.n} "
              f"ticks={r.ticks} finished={len(r.finish_order)} -> {'identical' if ok else 'DIFFERENT'}")
        if not ok:
            bad += 1
            keep = os.path.join(HERE, f"_fuzz_fail_{seed}")
            shutil.copytree(d, keep, dirs_exist_ok=True)
        shutil.rmtree(d)
    print(f"{n_cases} cases, {bad} different")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
