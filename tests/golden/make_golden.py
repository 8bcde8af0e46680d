#!/usr/bin/env python
"""Regenerate the golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py [case ...]

For every case it writes  tests/golden/<case>/{trace.csv,flags.json} (inputs)
and {job.csv,cluster.csv} (what /root/reference/run_sim.py produced).  The
reference is executed through runpy from a scratch CWD with
numpy.random.seed(SEED) set first, so that the one stochastic column
(avg_gpu_utilization, /root/reference/infra/device.py:52) is reproducible.
Nothing in the reference is edited; logging is silenced only by installing a
root handler before its basicConfig call runs.
"""
from __future__ import annotations

import glob
import json
import logging
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SEED = 7
sys.path.insert(0, REPO)

from gpuschedule_b200 import tracegen  # noqa: E402

HEADER = ["type", "normalized_time", "minutes", "gpu_per_container",
          "gpu_utilization_avg", "gpu_utilization_max", "memory_max",
          "memory_avg", "used_gpus"]


def kat0_frame():
    rows = [(0, 22, 37, 81, 2147483648, 3221225472, 4),
            (20000, 25, 70, 85, 5368709120, 3221225472, 4),
            (30000, 38, 48, 89, 6442450944, 3221225472, 8),
            (80000, 30, 12, 88, 3221225472, 2147483648, 16),
            (130000, 26, 63, 88, 4294967296, 3221225472, 1),
            (150000, 23, 61, 80, 7516192768, 1073741824, 16),
            (250000, 24, 69, 80, 6442450944, 2147483648, 4),
            (310000, 37, 22, 82, 7516192768, 3221225472, 2),
            (320000, 13, 16, 80, 3221225472, 3221225472, 8),
            (400000, 33, 70, 93, 5368709120, 2147483648, 4),
            (420000, 28, 11, 90, 7516192768, 2147483648, 2),
            (450000, 4, 47, 92, 5368709120, 1073741824, 2)]
    return pd.DataFrame([("noninteractive", nt, m, 1, ua, um, mm, ma, g)
                         for (nt, m, ua, um, mm, ma, g) in rows], columns=HEADER)


def ragged_frame():
    """Unsorted rows, ties, fractional times, NaNs and a filtered type."""
    df = tracegen.synth_frame(120, seed=11, rate=1.5)
    df = df.drop(columns=["model"])
    rng = np.random.default_rng(5)
    df["normalized_time"] = df["normalized_time"] + rng.integers(0, 4, size=len(df)) * 2500 + 1234567
    df.loc[rng.choice(len(df), 9, replace=False), "type"] = "interactive"
    df["minutes"] = df["minutes"].astype(float)
    df.loc[rng.choice(len(df), 5, replace=False), "minutes"] = np.nan
    df = df.sample(frac=1.0, random_state=3)          # shuffled, keeps labels
    return df


def leak_frame():
    """A job whose memory_max can never pass the 500 MiB device margin
    (/root/reference/infra/device.py:75) blocks the head of the queue and leaks
    cpu/mem on every node it probes (/root/reference/infra/node.py:204-221)."""
    df = tracegen.synth_frame(40, seed=21, rate=0.8).drop(columns=["model"])
    df.loc[6, "memory_max"] = 33000 * (1 << 20)     # >= (32768-500) MiB
    df.loc[6, "used_gpus"] = 2
    df.loc[17, "memory_max"] = 32300 * (1 << 20)    # cross-node variant
    df.loc[17, "used_gpus"] = 16
    return df


def floatgpu_frame():
    df = tracegen.synth_frame(50, seed=31, rate=0.7).drop(columns=["model"])
    df["used_gpus"] = df["used_gpus"].astype(float)
    return df


def netcost_frame(n, seed, rate):
    """Wide jobs (many span several nodes) with model_name / iterations / ps_count columns; iterations kept
    small so that the added seconds stay comparable with the durations."""
    df = tracegen.synth_frame(n, seed=seed, rate=rate, with_network=True,
                              gpu_choices=[1, 2, 4, 8, 12, 16, 24], gpu_probs=[.2, .15, .15, .2, .1, .1, .1])
    df = df.drop(columns=["model"])
    rng = np.random.default_rng(seed)
    df["iterations"] = rng.choice([1, 3, 8, 20, 57], size=n)
    df.loc[rng.choice(n, n // 5, replace=False), "ps_count"] = 1          # not distributed: no cost
    return df


def gpc2_frame():
    df = tracegen.synth_frame(200, seed=41, rate=1.0, gpu_per_container=2,
                              gpu_choices=[2, 4, 8, 16, 24], gpu_probs=[.3, .3, .2, .1, .1])
    return df.drop(columns=["model"])


CASES = {
    # name: (frame builder, flags)
    "kat0": (kat0_frame, dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)),
    "n64": (lambda: tracegen.synth_frame(64, seed=1), dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)),
    "sat300": (lambda: tracegen.synth_frame(300, seed=2), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8)),
    "n1000": (lambda: tracegen.synth_frame(1000, seed=1), dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)),
    "burst": (lambda: tracegen.synth_frame(400, seed=3, rate=3.0), dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8)),
    "gpc2": (gpc2_frame, dict(num_switch=2, num_node_p_switch=8, num_gpu_p_node=8)),
    "spec": (lambda: tracegen.synth_frame(150, seed=4, rate=0.4), dict(cluster_spec="cluster_spec.csv",
             _spec_row=dict(num_switch=1, num_node_p_switch=16, num_gpu_p_node=4, num_cpu_p_node=128, mem_p_node=254))),
    "tight": (lambda: tracegen.synth_frame(250, seed=5, rate=1.0), dict(num_switch=2, num_node_p_switch=6, num_gpu_p_node=8,
              num_cpu_p_node=60, mem_p_node=300)),
    "ragged": (ragged_frame, dict(num_switch=1, num_node_p_switch=8, num_gpu_p_node=8)),
    "leak": (leak_frame, dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8)),
    "floatgpu": (floatgpu_frame, dict(num_switch=1, num_node_p_switch=16, num_gpu_p_node=8)),
    # network-cost branch (schedule.py:48-53): the live Job lacks ps_count / model_size / iterations, so the run
    # only gets there with those three attributes attached (see _NETCOST_INJECT); everything else is unmodified
    "netcost": (lambda: netcost_frame(260, 51, 0.6), dict(num_switch=2, num_node_p_switch=8, num_gpu_p_node=8,
                enable_network_costs=True)),
    "netcost_lat": (lambda: netcost_frame(200, 52, 1.2), dict(num_switch=1, num_node_p_switch=12, num_gpu_p_node=4,
                    enable_network_costs=True, bandwidth=800, internode_latency=0.004)),
}

# Attaches the attributes network_service.calculate_network_costs reads (network_service.py:12,34,36) to every
# Job the reference builds: ps_count and iterations from the trace's extra columns, model_size from the
# reference's OWN table (model/model_factory.py:19-55) keyed by the trace's model_name column.
_NETCOST_INJECT = (
    "import importlib.abc, importlib.util\n"
    "import pandas as _pd\n"
    "from model import model_factory as _mf\n"
    "_t = _pd.read_csv('trace.csv')\n"
    "_extra = {str(i): (int(r.ps_count), _mf.model_sizes[r.model_name], r.iterations) for i, r in _t.iterrows()}\n"
    "def _patch(mod):\n"
    "    _init = mod.Job.__init__\n"
    "    def _patched(self, job_id, *a, **k):\n"
    "        _init(self, job_id, *a, **k)\n"
    "        self.ps_count, self.model_size, self.iterations = _extra[str(job_id)]\n"
    "    mod.Job.__init__ = _patched\n"
    "class _Hook(importlib.abc.MetaPathFinder):\n"        # core.jobs.job can only be imported once run_sim.py
    "    def find_spec(self, name, path, target=None):\n"  # has initialised base_factory: patch right after
    "        if name != 'core.jobs.job':\n"
    "            return None\n"
    "        sys.meta_path.remove(self)\n"
    "        spec = importlib.util.find_spec(name)\n"
    "        run = spec.loader.exec_module\n"
    "        def exec_module(mod):\n"
    "            run(mod)\n"
    "            _patch(mod)\n"
    "        spec.loader.exec_module = exec_module\n"
    "        return spec\n"
    "sys.meta_path.insert(0, _Hook())\n")


def run_reference(trace_csv: str, flags: dict, out_dir: str):
    scratch = tempfile.mkdtemp(prefix="gsref_")
    shutil.copy(trace_csv, os.path.join(scratch, "trace.csv"))
    argv = [os.path.join(REF, "run_sim.py"), "--scheme", flags.get("_scheme", "yarn"), "--schedule", flags.get("_schedule", "fifo"),
            "--trace_file", "trace.csv", "--log_path", "g"]
    for k, v in flags.items():
        if k.startswith("_"):
            continue
        if k == "cluster_spec":
            spec = os.path.join(scratch, "cluster_spec.csv")
            row = flags["_spec_row"]
            with open(spec, "w") as f:
                f.write(",".join(row.keys()) + "\n" + ",".join(str(x) for x in row.values()) + "\n")
            v = spec           # absolute: the reference opens it relative to its own root
        argv += ["--" + k, str(v)]
    code = (
        "import sys, os, runpy, logging, numpy\n"
        "logging.getLogger().addHandler(logging.NullHandler())\n"
        f"sys.path.insert(0, {REF!r}); os.chdir({scratch!r}); sys.argv = {argv!r}\n"
        + (_NETCOST_INJECT if flags.get("enable_network_costs") else "") +
        f"numpy.random.seed({SEED})\n"
        "try:\n"
        f"    runpy.run_path({os.path.join(REF, 'run_sim.py')!r}, run_name='__main__')\n"
        "except SystemExit:\n    pass\n")
    import subprocess
    subprocess.run([sys.executable, "-W", "ignore", "-c", code], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    runs = glob.glob(os.path.join(scratch, "log", "g", "*"))
    assert len(runs) == 1, runs
    for name in ("job.csv", "cluster.csv"):
        shutil.copy(os.path.join(runs[0], name), os.path.join(out_dir, name))
    shutil.rmtree(scratch)


def main(which):
    for name in which:
        build, flags = CASES[name]
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        df = build()
        trace = os.path.join(d, "trace.csv")
        df.to_csv(trace, index=False)
        meta = {"flags": {k: v for k, v in flags.items()}, "numpy_seed": SEED,
                "reference": "matthewygf/GPUSchedule @ ea0f1474, run_sim.py --scheme yarn --schedule fifo"}
        with open(os.path.join(d, "flags.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        run_reference(trace, flags, d)
        n_rows = sum(1 for _ in open(os.path.join(d, "cluster.csv"))) - 1
        n_jobs = sum(1 for _ in open(os.path.join(d, "job.csv"))) - 1
        print(f"{name}: ticks={n_rows} finished={n_jobs}")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
