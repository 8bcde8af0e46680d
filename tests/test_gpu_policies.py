"""Event-driven policies (sjf / dlas / dlas-gpu / gittins): CUDA engine vs the CPU restatement
(oracle/policy_oracle.c) and vs tests/golden/policy_*: the reference holds these policies as dead
code; the fixtures were produced by executing that dead loop code verbatim under
the stub harness of tests/golden/make_policy_golden.py, which is what pins both sides."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _policy(name, table):
    from gpuschedule_b200 import capi, policies
    if name == "sjf":
        return capi.make_policy("sjf")
    if name == "dlas":
        return capi.make_policy("dlas", num_queue=3, queue_limit=[40, 160])
    if name == "dlas-gpu":
        return capi.make_policy("dlas-gpu", num_queue=4, queue_limit=[300, 900, 3000])
    return capi.make_policy("gittins", gittins_delta=3250.0,
                            gittins_table=policies.build_gittins_table(policies.gittins_samples(table), 3250.0))


def _run(cluster, policy, table, rows_cap=0, engine=0):
    from gpuschedule_b200 import capi
    with capi.Engine(device=0, nsims=1) as eng:
        eng.set_engine(engine)            # 0: warp-cooperative kernels, 2: thread-per-replica fallback
        eng.config(0, cluster, policy)
        eng.load_trace(0, table)
        rows = eng.run_all(rows_cap=rows_cap)[0]
        recs, order = eng.fetch_jobs(0)
        return rows, recs, order, eng.stats(0)


CASES = [(800, 11, 0.5, dict(num_switch=1, num_node_p_switch=8)),      # saturated: preemptions
         (1500, 12, 0.5, dict(num_switch=4, num_node_p_switch=32)),    # BASELINE cluster, light load
         (600, 13, 2.0, dict(num_switch=1, num_node_p_switch=3, num_gpu_p_node=4))]


@pytest.mark.parametrize("name", ["sjf", "dlas", "dlas-gpu", "gittins"])
@pytest.mark.parametrize("cfg", CASES, ids=[f"seed{c[1]}" for c in CASES])
def test_policy_engine_matches_oracle(name, cfg):
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    n, seed, rate, ckw = cfg
    cluster = capi.make_cluster(**ckw)
    table = ingest.table_from_columns(tracegen.synth_columns(n, seed=seed, rate=rate,
                                                             gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1]))
    pol = _policy(name, table)
    ref = oracle.run_policy(cluster, pol, table)
    rows, recs, order, st = _run(cluster, pol, table)
    assert st.ticks == ref.ticks and st.done == 1
    assert np.array_equal(order, ref.finish_order)
    assert rows.tobytes() == ref.rows.tobytes()
    assert recs.tobytes() == ref.recs.tobytes()
    assert st.events == ref.events
    # window resume gives the same answer
    rows2, recs2, order2, st2 = _run(cluster, pol, table, rows_cap=53)
    assert rows2.tobytes() == ref.rows.tobytes() and recs2.tobytes() == ref.recs.tobytes()
    # and so does the thread-per-replica fallback kernel
    rows3, recs3, order3, st3 = _run(cluster, pol, table, engine=2)
    assert rows3.tobytes() == ref.rows.tobytes() and recs3.tobytes() == ref.recs.tobytes() and st3.events == ref.events


def test_policy_invariants():
    """Properties that hold for every policy: conservation, monotone time, preemption accounting."""
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=1, num_node_p_switch=8)
    table = ingest.table_from_columns(tracegen.synth_columns(3000, seed=21, rate=0.6,
                                                             gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1]))
    need = np.maximum(1, np.ceil(table.duration)).astype(np.int32)
    for name in ["sjf", "dlas-gpu", "gittins"]:
        rows, recs, order, st = _run(cluster, _policy(name, table), table)
        assert st.finished == table.n and sorted(order.tolist()) == list(range(table.n))
        assert np.all(np.diff(rows["now"]) >= 0)
        assert np.all(rows["busy_gpus"] <= 64) and np.all(rows["busy_gpus"] + rows["idle_gpus"] == 64)
        assert np.all(recs["start"] >= table.arrive_tick) and np.all(recs["end"] - recs["start"] >= need)
        assert np.all(recs["jct"] == need) and np.all(recs["preempt"] >= 1)
        assert rows["finished"][-1] == table.n and rows["running"][-1] == 0


def test_mixed_policies_in_one_handle():
    """fifo and event-driven replicas side by side in one launch (BASELINE config 5 style)."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=8)
    names = ["fifo", "sjf", "dlas-gpu", "gittins"] * 10
    tables = [ingest.table_from_columns(tracegen.synth_columns(300 + 7 * i, seed=300 + i, rate=0.8,
                                                               gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1]))
              for i in range(len(names))]
    with capi.Engine(device=0, nsims=len(names)) as eng:
        pols = []
        for i, nm in enumerate(names):
            pols.append(capi.make_policy("fifo") if nm == "fifo" else _policy(nm, tables[i]))
            eng.config(i, cluster, pols[i])
            eng.load_trace(i, tables[i])
        rows = eng.run_all()
        for i, nm in enumerate(names):
            ref = oracle.run_fifo(cluster, tables[i]) if nm == "fifo" else oracle.run_policy(cluster, pols[i], tables[i])
            recs, order = eng.fetch_jobs(i)
            assert rows[i].tobytes() == ref.rows.tobytes(), (i, nm)
            assert recs.tobytes() == ref.recs.tobytes(), (i, nm)
            assert np.array_equal(order, ref.finish_order), (i, nm)


def test_cli_runs_an_event_driven_policy(tmp_path):
    """run_sim.py --schedule dlas-gpu end to end: files written, rows/jobs equal the oracle's."""
    import glob
    import os
    import subprocess
    import sys
    import oracle
    from conftest import REPO
    from gpuschedule_b200 import capi, ingest, log_manager, tracegen
    df = tracegen.synth_frame(400, seed=31, rate=0.9, gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1])
    df.to_csv(tmp_path / "trace.csv", index=False)
    cmd = [sys.executable, os.path.join(REPO, "run_sim.py"), "--num_switch", "1", "--num_node_p_switch", "8",
           "--scheme", "yarn", "--schedule", "dlas-gpu", "--num_queue", "4", "--queue_limit", "300,900,3000",
           "--trace_file", "trace.csv", "--log_path", "p"]
    subprocess.run(cmd, cwd=tmp_path, check=True, capture_output=True)
    runs = glob.glob(str(tmp_path / "log" / "p" / "*"))
    assert len(runs) == 1
    table = ingest.JobTraceReader(str(tmp_path / "trace.csv")).prepare_jobs().table(0.5)
    cluster = capi.make_cluster(1, 8, 8)
    ref = oracle.run_policy(cluster, capi.make_policy("dlas-gpu", num_queue=4, queue_limit=[300, 900, 3000]), table)
    exp_jobs = log_manager.render_job_csv(table, ref.recs, ref.finish_order)
    assert open(os.path.join(runs[0], "job.csv"), newline="").read() == exp_jobs
    got_rows = open(os.path.join(runs[0], "cluster.csv"), newline="").read().split("\r\n")
    assert len(got_rows) == ref.ticks + 2 and got_rows[1].split(",")[0] == str(int(ref.rows["now"][0]))


def _fixture_cases():
    import os
    from conftest import GOLDEN
    return sorted(d for d in os.listdir(GOLDEN) if d.startswith("policy_") and os.path.isfile(os.path.join(GOLDEN, d, "expected.json")))


@pytest.mark.parametrize("engine", [0, 2], ids=["warp", "thread"])
@pytest.mark.parametrize("case", _fixture_cases())
def test_policy_engine_matches_reference_loop_fixtures(case, engine):
    """CUDA engine vs the outputs of the reference's own loop code (completions, checkpoints, event totals);
    policy_dlas_gpu_8q exercises the inherited-end-list quirk (Q25) and multi-level demotion jumps."""
    from types import SimpleNamespace
    from test_policy_golden import _load, check_against_expected
    table, cluster, pol, exp, kw = _load(case)
    rows, recs, order, st = _run(cluster, pol, table, engine=engine)
    assert st.done == 1
    check_against_expected(table, SimpleNamespace(finish_order=order, recs=recs, rows=rows, ticks=st.ticks, events=st.events), exp)


@pytest.mark.parametrize("ckw,rate,seed", [(dict(num_switch=2, num_node_p_switch=20, num_gpu_p_node=8), 6.0, 21),
                                           (dict(num_switch=1, num_node_p_switch=5, num_gpu_p_node=4), 3.0, 22),
                                           (dict(num_switch=4, num_node_p_switch=32, num_gpu_p_node=8, num_cpu_p_node=16), 8.0, 23)])
def test_sjf_long_lists_mixed_tasks_and_oversize_jobs(ckw, rate, seed):
    """sjf places runs of identical list entries with one walk over the nodes: lists of hundreds of entries (several
    chunks, runs crossing chunk borders), 1 / 2 / 4 GPUs per task mixed inside one GPU count, jobs too large for a
    device in the middle of runs, cross-node jobs, a slot-bound cluster (16 cpus per node)"""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(**ckw)
    cols = tracegen.synth_columns(2500, seed=seed, rate=rate, gpu_choices=[1, 2, 4, 8, 16, 32], gpu_probs=[.3, .2, .2, .15, .1, .05],
                                  max_mem_mib=40000)               # some jobs exceed the 32 GiB device
    rng = np.random.default_rng(seed)
    gpc = rng.choice([1, 2, 4], size=2500, p=[.6, .25, .15])
    cols["gpu_per_container"] = np.minimum(gpc, cols["used_gpus"]).astype(np.int64)
    table = ingest.table_from_columns(cols)
    pol = capi.make_policy("sjf")
    ref = oracle.run_policy(cluster, pol, table)
    assert int(ref.rows["queued"].max()) > 100                      # several chunks of runnable jobs
    for engine in (0, 2):
        rows, recs, order, st = _run(cluster, pol, table, engine=engine)
        assert rows.tobytes() == ref.rows.tobytes() and recs.tobytes() == ref.recs.tobytes(), engine
        assert np.array_equal(order, ref.finish_order) and st.events == ref.events, engine
    rows2, recs2, _, _ = _run(cluster, pol, table, rows_cap=97)
    assert rows2.tobytes() == ref.rows.tobytes() and recs2.tobytes() == ref.recs.tobytes()


@pytest.mark.parametrize("stretch", [1, 40000], ids=["direct-table", "bisection"])
def test_gittins_lookup_forms_agree_with_the_oracle(stretch):
    """gs_config_sim tabulates the gittins index per whole unit of attained service when the largest sample allows it and
    the kernel then reads it with one load; a table whose samples are far apart (here: stretched by 40000) keeps the
    bisection.  Both against oracle/policy_oracle.c, which always bisects."""
    import oracle
    from gpuschedule_b200 import capi, ingest, policies, tracegen
    cluster = capi.make_cluster(num_switch=1, num_node_p_switch=6)
    table = ingest.table_from_columns(tracegen.synth_columns(1200, seed=31, rate=1.5, gpu_choices=[1, 2, 4, 8], gpu_probs=[.4, .3, .2, .1]))
    samples = policies.gittins_samples(table) * stretch
    data, index = policies.build_gittins_table(samples, 3250.0 * stretch)
    if stretch > 1:                                         # the same index values at far-apart, partly non-integer sample positions
        base = np.sort(policies.gittins_samples(table)).astype(np.float64) + 0.25 * (np.arange(len(data) - 1) % 2)
        data = np.concatenate([np.sort(base) * stretch, data[-1:]])
    pol = capi.make_policy("gittins", gittins_delta=3250.0, gittins_table=(data, index))
    ref = oracle.run_policy(cluster, pol, table)
    for engine in (0, 2):
        rows, recs, order, st = _run(cluster, pol, table, engine=engine)
        assert rows.tobytes() == ref.rows.tobytes() and recs.tobytes() == ref.recs.tobytes(), engine
        assert np.array_equal(order, ref.finish_order) and st.events == ref.events, engine
    assert st.events > 3 * 1200                             # preemptions happened: ranks were compared, not only read
