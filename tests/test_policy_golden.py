"""Pins the event-driven policy restatement (oracle/policy_oracle.c) against the reference's own loop
code: tests/golden/policy_*/expected.json were produced by executing smallest_first_sim_jobs /
dlas_sim_jobs / gittins_sim_jobs / parse_job_dist of /root/reference/run_sim.py VERBATIM under the
stub harness of tests/golden/make_policy_golden.py (the stubs supply only what the dead code leaves
undefined).  Compared: every completion (job, time, first start, number of (re)starts) in order, and
every checkpoint (time, running, pending, busy GPUs, total pending time) -- i.e. event selection,
aging, demotion, ordering, admission, preemption and jump logic all follow the reference's code."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _cases():
    return sorted(d for d in os.listdir(GOLDEN) if d.startswith("policy_") and os.path.isfile(os.path.join(GOLDEN, d, "expected.json")))


def _load(case):
    from gpuschedule_b200 import capi, ingest, policies
    d = os.path.join(GOLDEN, case)
    meta = json.load(open(os.path.join(d, "params.json")))
    exp = json.load(open(os.path.join(d, "expected.json")))
    table = ingest.JobTraceReader(os.path.join(d, "trace.csv")).prepare_jobs().table(0.5)
    cluster = capi.make_cluster(**meta["cluster"])
    kw = dict(meta["params"])
    if meta["policy"] == "gittins":
        kw["gittins_table"] = policies.build_gittins_table(policies.gittins_samples(table), kw.get("gittins_delta", 3250.0))
    return table, cluster, capi.make_policy(meta["policy"], **kw), exp, kw


def check_against_expected(table, res, exp):
    comp = exp["completions"]
    assert len(res.finish_order) == len(comp)
    got = [[int(j), int(res.recs["end"][j]), int(res.recs["start"][j]), int(res.recs["preempt"][j])] for j in res.finish_order]
    assert got == [[c[0], c[1], c[2], c[3]] for c in comp]
    chk = exp["checkpoints"]
    assert res.ticks == len(chk)
    rows = res.rows
    got_rows = np.stack([rows["now"], rows["running"], rows["queued"], rows["busy_gpus"], rows["pend_sum"]], axis=1).tolist()
    assert got_rows == chk
    # preemptions: every RUNNING->PENDING flip is an event; total from the reference's job['preempt'] counters
    left = exp["unfinished"]                       # jobs still runnable when the reference loop gave up
    assert res.events == (len(comp) + len(left)) + len(comp) + sum(c[3] + c[4] for c in comp) + sum(u[1] + u[2] for u in left)


@pytest.mark.parametrize("case", _cases())
def test_policy_oracle_matches_reference_loops(case):
    import oracle
    table, cluster, pol, exp, kw = _load(case)
    res = oracle.run_policy(cluster, pol, table)
    check_against_expected(table, res, exp)


def test_gittins_table_matches_reference_parse_job_dist():
    """policies.build_gittins_table == the reference's parse_job_dist + cal_r_gittins_index (executed verbatim)."""
    from gpuschedule_b200 import policies
    table, cluster, pol, exp, kw = _load("policy_gittins")
    data, gi = policies.build_gittins_table(policies.gittins_samples(table), 3250.0)
    assert data.tolist() == exp["gittins_table"]["data"]
    assert gi.tolist() == exp["gittins_table"]["gittins"]


def _gittins_table_scalar(samples, delta):
    """cal_r_gittins_index (run_sim.py:1650-1708) value by value with Python's round(): what the vectorised builder must equal"""
    import bisect
    import sys
    data = sorted(int(x) for x in samples)
    num = len(data)
    prefix = [0]
    for v in data:
        prefix.append(prefix[-1] + v)
    last = data[-1]

    def r_index(a):
        if a > last - 1:
            return 0.0
        idx = bisect.bisect_right(data, a)
        next_a = a + delta
        idx_delta = num - 1 if next_a > last - 1 else bisect.bisect_right(data, next_a)
        p = round(((idx_delta - idx) * 1.0) / (num - idx), 5)
        e = round(((prefix[idx_delta] - prefix[idx]) + (delta * (num - idx_delta))) / (num - idx), 5)
        return round(p * 1000000 / e, 4)
    return (np.array([float(v) for v in data] + [float(sys.maxsize)]), np.array([r_index(int(v - 1)) for v in data] + [0.0]))


def test_vectorised_gittins_table_equals_the_scalar_restatement_bit_for_bit():
    from gpuschedule_b200 import policies, tracegen, ingest
    rng = np.random.default_rng(3)
    cases = [policies.gittins_samples(ingest.table_from_columns(tracegen.synth_columns(20000, seed=s))) for s in (1, 2)]
    cases += [rng.integers(1, 50, 5000), rng.integers(1, 10 ** 7, 20000), np.array([5]), np.array([1, 1, 1, 2]),
              rng.integers(3000, 3500, 3000), np.arange(1, 4000)]
    for delta in (3250.0, 10.0, 123.5):
        for c in cases:
            a, b = policies.build_gittins_table(c, delta), _gittins_table_scalar(c, delta)
            assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), (delta, len(c))
    for nd in (4, 5):                                       # the rounding helper alone, half-way decimals included
        x = np.concatenate([rng.random(50000) * 10.0 ** rng.integers(-3, 7, 50000), (rng.integers(0, 10 ** 7, 50000) * 10 + 5) / 10.0 ** (nd + 1),
                            rng.integers(0, 10 ** 6, 20000) / 1024.0, np.array([0.0, 0.5, 1.5, 2.5e-5, 0.000005, 0.000015, 0.000025, 1e-7])])
        assert policies._py_round(x, nd).tobytes() == np.array([round(float(v), nd) for v in x]).tobytes()
