"""BASELINE.json configs at their stated sizes, every one compared with the CPU checkers on FULL structures (every row,
every job record, the finish order) -- not only on event totals:

  C2  10k-job trace, sjf + yarn, 4x32x8
  C3  100k-job trace, dlas-gpu (4-queue MLFQ, thresholds 3600/7200/18000 of the README), 4x32x8
  C4  100k-job trace, gittins index (delta 3250), 4x32x8   (one GPU here; two GPUs: tests/test_gpu_sharded.py)
  C5  16x64x8 cluster (8192 GPUs), mixed policies: fifo on a 1M-job trace, sjf / dlas-gpu / gittins on 100k / 20k jobs

The traces are the SURVEY 8(d) generator's (bench.fast_table); C5 arrives 8x faster (4 jobs per tick) because the cluster
is 8x larger.  At most one fifo job starts per tick, so the 1M-job fifo run is the deep-queue regime (a million ticks, a
queue of several hundred thousand jobs) -- the regime the BASELINE trace never reaches."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def _policy_case(name, n, cluster_kw, rate, seed):
    import bench
    import oracle
    from gpuschedule_b200 import capi
    cluster = capi.make_cluster(**cluster_kw)
    table = bench.fast_table(n, seed, rate=rate)
    pol = bench.make_policy(name, table)
    with capi.Engine(device=0, nsims=1) as eng:
        eng.config(0, cluster, pol)
        eng.load_trace(0, table)
        rows = eng.run_all()[0]
        recs, order = eng.fetch_jobs(0)
        st = eng.stats(0)
    ref = oracle.run_policy(cluster, pol, table)
    assert st.done == 1 and st.events == ref.events and len(rows) == ref.ticks, name
    assert rows.tobytes() == ref.rows.tobytes(), name
    assert recs.tobytes() == ref.recs.tobytes(), name
    assert np.array_equal(order, ref.finish_order), name
    return st


def test_c2_sjf_10k():
    st = _policy_case("sjf", 10000, dict(num_switch=4, num_node_p_switch=32), 0.5, 1)
    assert st.finished == 10000


def test_c3_dlas_gpu_100k_readme_thresholds():
    st = _policy_case("dlas-gpu", 100000, dict(num_switch=4, num_node_p_switch=32), 0.5, 1)
    assert st.finished == 100000


def test_c4_gittins_100k_one_gpu():
    st = _policy_case("gittins", 100000, dict(num_switch=4, num_node_p_switch=32), 0.5, 1)
    assert st.finished == 100000


@pytest.mark.parametrize("name,n", [("sjf", 20000), ("dlas-gpu", 100000), ("gittins", 100000)])
def test_c5_policies_on_16x64x8(name, n):
    _policy_case(name, n, dict(num_switch=16, num_node_p_switch=64), 4.0, 6)


def test_c3_like_saturated_policies_preempt_and_match():
    """the BASELINE arrival rate keeps 4x32x8 at ~10 % load, where the policies never preempt; 3 jobs per tick do"""
    for name in ("dlas-gpu", "gittins"):
        st = _policy_case(name, 30000, dict(num_switch=4, num_node_p_switch=32), 3.0, 7)
        assert st.events >= 3 * 30000


def test_c5_fifo_1m_jobs_16x64x8():
    """1M jobs on 8192 GPUs: the engine against oracle/tight2_cpu.c (itself pinned to the literal oracle on the fixtures
    and on random cases, tests/test_tight2_cpu.py) on every row, record, span and the finish order; the literal oracle
    re-scans 8192 devices per tick and needs ~20 minutes for this size, so it checks a 100k-job prefix-size run instead."""
    import bench
    import oracle
    from gpuschedule_b200 import capi
    cluster = capi.make_cluster(16, 64, 8)
    table = bench.fast_table(1000000, 5, rate=4.0)
    with capi.Engine(device=0, nsims=1) as eng:
        eng.set_span_budget(2.0)
        eng.config(0, cluster)
        eng.load_trace(0, table)
        rows = eng.run_all()[0]
        recs, order = eng.fetch_jobs(0)
        span_off, spans = eng.fetch_spans(0)
        st = eng.stats(0)
    ref = oracle.Tight2(cluster, table).run_all()
    assert st.done == 1 and st.finished == 1000000 and st.ticks == ref.ticks and st.events == ref.events
    assert rows.tobytes() == ref.rows.tobytes()
    assert recs.tobytes() == ref.recs.tobytes() and np.array_equal(order, ref.finish_order)
    assert np.array_equal(span_off, ref.span_off) and spans.tobytes() == ref.spans.tobytes()
    assert st.placement_evals == ref.evals
    assert int(rows["queued"].max()) > 100000                         # the deep-queue regime
    assert len(np.unique(recs["start"])) == 1000000                    # one start per tick (Q1)


def test_c5_fifo_100k_against_the_literal_oracle():
    import bench
    import oracle
    from gpuschedule_b200 import capi
    cluster = capi.make_cluster(16, 64, 8)
    table = bench.fast_table(20000, 5, rate=4.0)
    with capi.Engine(device=0, nsims=1) as eng:
        eng.config(0, cluster)
        eng.load_trace(0, table)
        rows = eng.run_all()[0]
        recs, order = eng.fetch_jobs(0)
        st = eng.stats(0)
    ref = oracle.run_fifo(cluster, table)
    assert rows.tobytes() == ref.rows.tobytes() and recs.tobytes() == ref.recs.tobytes()
    assert np.array_equal(order, ref.finish_order) and st.placement_evals == ref.evals
