"""CPU checks of the event-driven policy restatement (oracle/policy_oracle.c) and the gittins
table builder: internal consistency; parity with the reference's loop code is tests/test_policy_golden.py."""
import numpy as np


def _table(n=600, seed=4, rate=0.7):
    from gpuschedule_b200 import ingest, tracegen
    return ingest.table_from_columns(tracegen.synth_columns(n, seed=seed, rate=rate,
                                                            gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1]))


def test_gittins_table_matches_literal_restatement():
    """The O(n log n) builder equals a literal transcription of run_sim.py:1650-1708 on a small sample."""
    import sys
    from gpuschedule_b200 import policies
    rng = np.random.default_rng(3)
    sample = sorted(int(x) for x in rng.integers(1, 20000, size=200))
    delta = 3250.0
    num = len(sample)

    def literal(a):
        if a > sample[-1] - 1:
            return 0.0
        idx = next(i for i, v in enumerate(sample) if v > a)
        nxt = a + delta
        idx_d = num - 1 if nxt > sample[-1] - 1 else next(i for i, v in enumerate(sample) if v > nxt)
        p = round(((idx_d - idx) * 1.0) / (num - idx), 5)
        e = round((sum(sample[idx:idx_d]) + delta * (num - idx_d)) / (num - idx), 5)
        return round(p * 1000000 / e, 4)

    data, gi = policies.build_gittins_table(sample, delta)
    assert len(data) == num + 1 and data[-1] == float(sys.maxsize) and gi[-1] == 0.0
    assert [literal(int(v - 1)) for v in sample] == gi[:-1].tolist()


def test_policy_oracle_conservation_and_work():
    import oracle
    from gpuschedule_b200 import capi, policies
    t = _table()
    cluster = capi.make_cluster(1, 8, 8)
    need = np.maximum(1, np.ceil(t.duration)).astype(np.int32)
    pols = {"sjf": capi.make_policy("sjf"),
            "dlas-gpu": capi.make_policy("dlas-gpu", num_queue=4, queue_limit=[300, 900, 3000]),
            "gittins": capi.make_policy("gittins", gittins_table=policies.build_gittins_table(policies.gittins_samples(t)))}
    for name, pol in pols.items():
        r = oracle.run_policy(cluster, pol, t)
        assert sorted(r.finish_order.tolist()) == list(range(t.n)), name
        assert np.all(r.recs["jct"] == need) and np.all(r.recs["end"] - r.recs["start"] >= need), name
        assert np.all(np.diff(r.rows["now"]) >= 0) and np.all(r.rows["busy_gpus"] <= 64), name
        # events = arrivals + completions + every (re)start + every preemption
        assert r.events == 2 * t.n + int(r.recs["preempt"].sum()) + int((r.recs["preempt"] - 1).sum()), name


def test_sjf_prefers_small_jobs_under_contention():
    """With a saturated cluster the mean wait of 1-GPU jobs under sjf is below that of 8-GPU jobs."""
    import oracle
    from gpuschedule_b200 import capi
    t = _table(n=1200, seed=9, rate=1.5)
    r = oracle.run_policy(capi.make_cluster(1, 4, 8), capi.make_policy("sjf"), t)
    wait = r.recs["start"] - t.arrive_tick
    assert wait[t.gpus == 1].mean() < wait[t.gpus == 8].mean()
