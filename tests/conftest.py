import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.isfile(os.path.join(GOLDEN, d, "flags.json")))


def load_golden(case):
    """(JobTable, GsCluster, meta, expected job.csv text, expected cluster.csv text)"""
    from gpuschedule_b200 import capi, ingest
    d = os.path.join(GOLDEN, case)
    with open(os.path.join(d, "flags.json")) as f:
        meta = json.load(f)
    flags = dict(meta["flags"])
    spec = flags.pop("_spec_row", None)
    flags.pop("cluster_spec", None)
    if spec:
        flags.update(spec)
    cluster = capi.make_cluster(**flags)
    reader = ingest.JobTraceReader(os.path.join(d, "trace.csv"))
    table = reader.prepare_jobs().table(0.5)
    with open(os.path.join(d, "job.csv"), newline="") as f:
        job_csv = f.read()
    with open(os.path.join(d, "cluster.csv"), newline="") as f:
        cluster_csv = f.read()
    return table, cluster, meta, job_csv, cluster_csv


def render_outputs(table, cluster, rows, recs, finish_order, span_off, spans, seed):
    """Format engine/oracle results exactly as the CLI does."""
    import numpy as np
    from gpuschedule_b200 import log_manager, rngcol
    m = cluster.num_switch * cluster.num_node_p_switch
    g = cluster.num_gpu_p_node
    np.random.seed(seed)
    util = rngcol.utilization_text(len(rows), m, g, table, recs, span_off, spans)
    cluster_csv = log_manager.render_cluster_csv(rows, util, m * g * cluster.gpu_mem_cap_mib)
    job_csv = log_manager.render_job_csv(table, recs, finish_order)
    return job_csv, cluster_csv


def horus_cases():
    return sorted(d for d in os.listdir(GOLDEN) if os.path.isfile(os.path.join(GOLDEN, d, "horus.json")))


def load_horus(case):
    """(JobTable, GsCluster, run parameters, expected job.csv text, expected cluster.csv text) of a horus_* fixture."""
    from gpuschedule_b200 import capi, ingest
    d = os.path.join(GOLDEN, case)
    with open(os.path.join(d, "horus.json")) as f:
        meta = json.load(f)
    flags = dict(meta["flags"])
    params = dict(scheme=flags.pop("_scheme"), schedule=flags.pop("_schedule"), num_buffer=flags.pop("num_buffer", 5),
                  num_queue=flags.pop("num_queue", 1), seed=meta["numpy_seed"])
    cluster = capi.make_cluster(**flags)
    table = ingest.JobTraceReader(os.path.join(d, "trace.csv")).prepare_jobs().table(0.5)
    with open(os.path.join(d, "job.csv"), newline="") as f:
        job_csv = f.read()
    with open(os.path.join(d, "cluster.csv"), newline="") as f:
        cluster_csv = f.read()
    return table, cluster, params, job_csv, cluster_csv


def render_horus_outputs(table, cluster, res):
    """job.csv / cluster.csv text of a horus-family result, through the package's own formatters."""
    from gpuschedule_b200 import log_manager, rngcol
    m = cluster.num_switch * cluster.num_node_p_switch
    g = cluster.num_gpu_p_node
    util = rngcol.sampled_utilization_text(res.util, res.util_is_array)
    cluster_csv = log_manager.render_cluster_csv(res.rows, util, m * g * cluster.gpu_mem_cap_mib)
    return log_manager.render_horus_job_csv(table, res.recs, res.finish_order), cluster_csv


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
