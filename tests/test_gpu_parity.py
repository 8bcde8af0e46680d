"""GPU parity: the CUDA engine (through the C ABI) against the reference-generated golden
fixtures and against the CPU oracle on seeded inputs.  Bit-exact everywhere (integer /
byte / index work; the one fp64 expression is checked for exact equality too)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, render_outputs

pytestmark = pytest.mark.gpu


ENGINES = [0]             # the fifo engine has one mapping (a warp per replica, event stepped); the parameter is kept so that
                          # the test ids stay comparable with round 1


def _engine_run(cluster, table, rows_cap=0, nsims=1, engine=0):
    from gpuschedule_b200 import capi
    with capi.Engine(device=0, nsims=nsims) as eng:
        eng.set_engine(engine)
        for s in range(nsims):
            eng.config(s, cluster)
            eng.load_trace(s, table)
        rows = eng.run_all(rows_cap=rows_cap)
        out = []
        for s in range(nsims):
            recs, order = eng.fetch_jobs(s)
            span_off, spans = eng.fetch_spans(s)
            out.append((rows[s], recs, order, span_off, spans, eng.stats(s)))
    return out


def _assert_same(ref, got, tag=""):
    rows, recs, order, span_off, spans, st = got
    assert st.ticks == ref.ticks, tag
    assert np.array_equal(order, ref.finish_order), tag
    if rows.tobytes() != ref.rows.tobytes():
        for i in range(min(len(rows), len(ref.rows))):
            assert rows[i].tobytes() == ref.rows[i].tobytes(), f"{tag} row {i}: {rows[i]} != {ref.rows[i]}"
    assert recs.tobytes() == ref.recs.tobytes(), tag
    assert np.array_equal(span_off, ref.span_off), tag
    assert spans.tobytes() == ref.spans.tobytes(), tag
    assert st.events == ref.events, tag
    assert st.placement_evals == ref.evals, tag


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case", golden_cases())
def test_engine_matches_reference_bytes(case, engine):
    """job.csv + all 13 cluster.csv columns byte-identical to the unmodified reference."""
    table, cluster, meta, job_csv, cluster_csv = load_golden(case)
    rows, recs, order, span_off, spans, st = _engine_run(cluster, table, engine=engine)[0]
    got_job, got_cluster = render_outputs(table, cluster, rows, recs, order, span_off, spans,
                                          meta["numpy_seed"])
    assert got_job == job_csv
    assert got_cluster == cluster_csv


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case", golden_cases())
def test_engine_matches_oracle_structs(case, engine):
    import oracle
    table, cluster, _, _, _ = load_golden(case)
    _assert_same(oracle.run_fifo(cluster, table), _engine_run(cluster, table, engine=engine)[0], case)


SEEDED = [
    # (jobs, seed, rate, cluster kwargs, tracegen kwargs)
    (5000, 101, 0.5, dict(num_switch=4, num_node_p_switch=32), {}),
    (3000, 102, 2.0, dict(num_switch=4, num_node_p_switch=32), {}),                  # saturating
    (1500, 103, 1.0, dict(num_switch=1, num_node_p_switch=5), {}),                   # M < 32, heavy queue
    (2000, 104, 0.7, dict(num_switch=3, num_node_p_switch=23, num_gpu_p_node=4), {}),  # M=69, G=4
    (2000, 105, 0.6, dict(num_switch=16, num_node_p_switch=64), {}),                 # M=1024
    (2000, 106, 1.0, dict(num_switch=2, num_node_p_switch=20, num_gpu_p_node=16, num_cpu_p_node=100, mem_p_node=400),
     dict(gpu_per_container=2, gpu_choices=[2, 4, 8, 16, 32, 64], gpu_probs=[.3, .2, .2, .15, .1, .05])),
    (1500, 107, 0.9, dict(num_switch=2, num_node_p_switch=16, gpu_memory_capacity=16), dict(max_mem_mib=17000)),  # leaks
]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("cfg", SEEDED, ids=[f"seed{c[1]}" for c in SEEDED])
def test_engine_matches_oracle_seeded(cfg, engine):
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    n, seed, rate, ckw, tkw = cfg
    cluster = capi.make_cluster(**ckw)
    table = ingest.table_from_columns(tracegen.synth_columns(n, seed=seed, rate=rate, **tkw))
    _assert_same(oracle.run_fifo(cluster, table), _engine_run(cluster, table, engine=engine)[0], f"seed{seed}")


def test_row_window_resume_is_identical():
    """A tiny device row window forces many launches; state persists across them."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=9)
    table = ingest.table_from_columns(tracegen.synth_columns(800, seed=7, rate=1.2))
    ref = oracle.run_fifo(cluster, table)
    for engine in ENGINES:
        for cap in (1, 7, 64, 1000):
            _assert_same(ref, _engine_run(cluster, table, rows_cap=cap, engine=engine)[0], f"rows_cap={cap}")


def test_short_row_windows_with_fetch_between_launches():
    """Rows are fetched window by window (97 records per launch); every window is self-contained."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=9)
    table = ingest.table_from_columns(tracegen.synth_columns(600, seed=8, rate=1.1))
    ref = oracle.run_fifo(cluster, table)
    with capi.Engine(device=0, nsims=1) as eng:
        eng.config(0, cluster)
        eng.load_trace(0, table)
        parts, seen, k = [], 0, 0
        while True:
            k += 1
            eng.run(0 if k % 3 else 40, 97)        # every third launch is also limited to 40 ticks
            st = eng.stats(0)
            parts.append(eng.fetch_rows(0, seen, st.ticks - seen))
            seen = st.ticks
            if st.done:
                break
        recs, order = eng.fetch_jobs(0)
        span_off, spans = eng.fetch_spans(0)
        _assert_same(ref, (np.concatenate(parts), recs, order, span_off, spans, eng.stats(0)), "alternate")


@pytest.mark.parametrize("engine", ENGINES)
def test_replicas_are_independent(engine):
    """Many replicas in one launch, different traces and clusters."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    nsims = 150
    tables, clusters = [], []
    for s in range(nsims):
        clusters.append(capi.make_cluster(num_switch=1 + s % 4, num_node_p_switch=8 + 3 * (s % 7)))
        tables.append(ingest.table_from_columns(tracegen.synth_columns(200 + 5 * s, seed=1000 + s, rate=0.4 + 0.05 * (s % 9))))
    with capi.Engine(device=0, nsims=nsims) as eng:
        eng.set_engine(engine)
        for s in range(nsims):
            eng.config(s, clusters[s])
            eng.load_trace(s, tables[s])
        rows = eng.run_all()
        for s in range(nsims):
            ref = oracle.run_fifo(clusters[s], tables[s])
            recs, order = eng.fetch_jobs(s)
            span_off, spans = eng.fetch_spans(s)
            _assert_same(ref, (rows[s], recs, order, span_off, spans, eng.stats(s)), f"replica {s}")


def test_network_cost_run_matches_oracle():
    """enable_network_costs: fp64 transfer time fused into the placement commit.  Exact
    equality of the resulting doubles (same operation order, no FMA contraction)."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=16, enable_network_costs=True)
    table = ingest.table_from_columns(tracegen.synth_columns(1500, seed=55, rate=0.3, with_network=True))
    assert table.model_mb is not None and table.ps_count is not None
    ref = oracle.run_fifo(cluster, table)
    assert np.any(ref.recs["duration"] != table.duration)       # the cost term is exercised
    for engine in ENGINES:
        _assert_same(ref, _engine_run(cluster, table, engine=engine)[0], "netcost")


def test_place_batch_matches_oracle():
    import oracle
    from gpuschedule_b200 import capi
    rng = np.random.default_rng(17)
    with capi.Engine(device=0, nsims=1) as eng:
        for m, g, cpu, mem in [(128, 8, 128, 512), (37, 4, 60, 300), (1024, 8, 128, 512), (5, 16, 400, 2000)]:
            cluster = capi.make_cluster(1, m, g, cpu, mem)
            nodes = np.zeros(m, dtype=capi.NODE_DTYPE)
            for i in range(m):
                k = int(rng.integers(0, g + 1))
                devs = rng.choice(g, size=k, replace=False)
                nodes["busy_mask"][i] = sum(1 << int(d) for d in devs)
                extra = int(rng.integers(0, 3))
                nodes["cpu_used"][i] = 12 * (k + extra)
                nodes["mem_used"][i] = 60 * (k + extra)
            b = 400
            jobs = np.zeros(b, dtype=capi.JOBREQ_DTYPE)
            gpc = rng.choice([1, 1, 1, 2, 4], size=b)
            jobs["gpu_per_task"] = gpc
            jobs["gpus"] = gpc * rng.integers(1, 40, size=b)
            jobs["mem_bytes"] = rng.integers(512, 34000, size=b).astype(np.int64) << 20
            task_off = np.zeros(b + 1, dtype=np.int64)
            np.cumsum(jobs["gpus"] // jobs["gpu_per_task"], out=task_off[1:])
            first, used, task_node, _ = eng.place_batch(cluster, nodes, jobs, task_off)
            for i in range(b):
                ok, f, u, tn = oracle.place_one(cluster, nodes, jobs[i])
                assert (first[i] >= 0) == ok, (m, i)
                assert first[i] == f and used[i] == u, (m, i, first[i], f, used[i], u)
                assert np.array_equal(task_node[task_off[i]:task_off[i + 1]], tn), (m, i)


def test_net_cost_matches_oracle():
    import oracle
    from gpuschedule_b200 import capi
    rng = np.random.default_rng(23)
    cluster = capi.make_cluster(4, 32, bandwidth=1250, internode_latency=0.015)
    b = 300
    sizes = rng.integers(1, 40, size=b)
    task_off = np.zeros(b + 1, dtype=np.int64)
    np.cumsum(sizes, out=task_off[1:])
    task_node = rng.integers(0, 6, size=int(task_off[-1])).astype(np.int32)
    is_ps = (rng.random(int(task_off[-1])) < 0.3).astype(np.uint8)
    ps_count = rng.integers(0, 4, size=b).astype(np.int32)
    model = rng.choice([15.0, 97.0, 233.0, 1300.0, 549.0], size=b)
    iters = rng.choice([1.0, 109.0, 521.0, 4861.0], size=b)
    with capi.Engine(device=0, nsims=1) as eng:
        for marks in (is_ps, None):
            got = eng.net_cost(cluster, task_off, task_node, marks, ps_count, model, iters)
            for i in range(b):
                seg = slice(task_off[i], task_off[i + 1])
                exp = oracle.net_cost(cluster, task_node[seg], None if marks is None else marks[seg],
                                      ps_count[i], model[i], iters[i])
                assert got[i] == exp, (i, got[i], exp)


def test_net_cost_matches_reference_function():
    """gs_net_cost == the reference's calculate_network_costs on the committed vectors (bit-exact doubles)."""
    import json
    import os
    from conftest import GOLDEN
    from gpuschedule_b200 import capi
    cases = json.load(open(os.path.join(GOLDEN, "netcost.json")))["cases"]
    groups = {}
    for c in cases:
        groups.setdefault((c["bandwidth"], c["latency"]), []).append(c)
    with capi.Engine(device=0, nsims=1) as eng:
        for (bw, lat), cs in groups.items():
            cluster = capi.make_cluster(4, 32, bandwidth=bw, internode_latency=lat)
            task_off = np.zeros(len(cs) + 1, dtype=np.int64)
            np.cumsum([len(c["node"]) for c in cs], out=task_off[1:])
            task_node = np.concatenate([np.array(c["node"], dtype=np.int32) for c in cs])
            marks = np.concatenate([np.array(c["is_ps"], dtype=np.uint8) for c in cs])
            got = eng.net_cost(cluster, task_off, task_node, marks, np.array([c["ps_count"] for c in cs], dtype=np.int32),
                               np.array([c["model_mb"] for c in cs]), np.array([c["iterations"] for c in cs]))
            for g, c in zip(got, cs):
                assert g == float.fromhex(c["expected"]), c


def test_full_size_properties_100k():
    """BASELINE size (100k jobs, 4x32x8): size-independent invariants + oracle equality."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(4, 32, 8)
    table = ingest.table_from_columns(tracegen.synth_columns(100000, seed=1, rate=0.5))
    rows, recs, order, span_off, spans, st = _engine_run(cluster, table)[0]
    n = table.n
    assert st.done == 1 and st.finished == n and st.events == 3 * n
    assert np.array_equal(rows["now"], np.arange(1, len(rows) + 1))
    assert np.all(rows["busy_gpus"] + rows["idle_gpus"] == 1024)
    assert np.all(rows["idle_nodes"] + rows["busy_nodes"] == 128)
    assert np.all(np.diff(rows["idle_nodes"]) <= 0) and np.all(np.diff(rows["finished"]) >= 0)
    assert rows["running"][-1] == 0 and rows["finished"][-1] == n
    assert np.all(recs["end"] - recs["start"] == recs["jct"])
    assert np.all(recs["jct"] == np.maximum(1, np.ceil(table.duration)).astype(np.int32))
    assert np.all(recs["start"] >= table.arrive_tick)
    assert len(np.unique(recs["start"])) == n                     # at most one start per tick (Q1)
    assert sorted(order.tolist()) == list(range(n))
    ends = recs["end"][order]
    assert np.all(np.diff(ends) >= 0)                              # finish order
    # per-span device masks are disjoint per node at any time: check total gpu count
    assert np.all(np.diff(span_off) >= 1)
    pc = np.array([bin(int(x)).count("1") for x in spans["devmask"]])
    assert np.array_equal(np.add.reduceat(pc, span_off[:-1]), table.gpus)
    ref = oracle.run_fifo(cluster, table)
    _assert_same(ref, (rows, recs, order, span_off, spans, st), "100k")


def test_cli_run_sim_writes_reference_bytes(tmp_path):
    """run_sim.py end to end: same flags, same files, same bytes as the reference run."""
    import glob
    import os
    import shutil
    import subprocess
    import sys
    from conftest import GOLDEN, REPO
    case = "n64"
    shutil.copy(os.path.join(GOLDEN, case, "trace.csv"), tmp_path / "trace.csv")
    cmd = [sys.executable, os.path.join(REPO, "run_sim.py"), "--num_switch", "4", "--num_node_p_switch", "32",
           "--num_gpu_p_node", "8", "--scheme", "yarn", "--schedule", "fifo", "--trace_file", "trace.csv",
           "--log_path", "g", "--enable_network_costs", "False", "--seed", "7"]
    subprocess.run(cmd, cwd=tmp_path, check=True, capture_output=True)
    runs = glob.glob(str(tmp_path / "log" / "g" / "*"))
    assert len(runs) == 1
    for name in ("job.csv", "cluster.csv"):
        got = open(os.path.join(runs[0], name), newline="").read()
        exp = open(os.path.join(GOLDEN, case, name), newline="").read()
        assert got == exp, name
    for name in ("cpu.csv", "gpu.csv", "memory.csv", "network.csv", "output.log"):
        assert os.path.exists(os.path.join(runs[0], name)), name


def test_packed_loader_and_fetch_all_match_column_api():
    """gs_load_trace_packed + gs_fetch_all are the same path with fewer host passes / syncs."""
    from gpuschedule_b200 import capi, ingest, tracegen
    from gpuschedule_b200.log_manager import JOB_DTYPE, ROW_DTYPE, SPAN_DTYPE
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=12)
    table = ingest.table_from_columns(tracegen.synth_columns(1200, seed=77, rate=0.9))
    base = _engine_run(cluster, table)[0]
    with capi.Engine(device=0, nsims=1) as eng:
        eng.config(0, cluster)
        eng.load_trace_packed(0, table.packed())
        eng.run(0, 0)
        st = eng.stats(0)
        assert st.done == 1
        rows = np.empty(st.ticks, dtype=ROW_DTYPE)
        jobs = np.empty(table.n, dtype=JOB_DTYPE)
        order = np.empty(table.n, dtype=np.int32)
        off = np.empty(table.n + 1, dtype=np.int64)
        spans = np.empty(len(base[4]) + 8, dtype=SPAN_DTYPE)
        r, j, o, so, sp = eng.fetch_all(0, rows, jobs, order, off, spans)
    assert r.tobytes() == base[0].tobytes() and j.tobytes() == base[1].tobytes()
    assert np.array_equal(o, base[2]) and np.array_equal(so, base[3]) and sp.tobytes() == base[4].tobytes()


EDGE = [
    # (name, cluster kwargs, frame mutation)
    ("g64_masks", dict(num_switch=1, num_node_p_switch=6, num_gpu_p_node=64, num_cpu_p_node=800, mem_p_node=4000), None),
    ("never_fits", dict(num_switch=1, num_node_p_switch=2, num_gpu_p_node=8), "huge"),
    ("single_job", dict(num_switch=1, num_node_p_switch=1, num_gpu_p_node=8), "one"),
]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case", EDGE, ids=[c[0] for c in EDGE])
def test_engine_edge_cases(case, engine):
    """64-GPU nodes (64-bit device masks), a job larger than the whole cluster (blocks the queue
    head for ever -> early exit, quirk Q4), and a one-job trace."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    name, ckw, mut = case
    cluster = capi.make_cluster(**ckw)
    if name == "g64_masks":
        cols = tracegen.synth_columns(500, seed=64, rate=0.8, gpu_choices=[1, 8, 48, 64, 128, 200], gpu_probs=[.3, .3, .15, .1, .1, .05])
    elif mut == "one":
        cols = tracegen.synth_columns(1, seed=65)
    else:
        cols = tracegen.synth_columns(60, seed=66, rate=1.0)
        cols["used_gpus"][7] = 64                      # 2 nodes x 8 GPUs can never host it
    table = ingest.table_from_columns(cols)
    ref = oracle.run_fifo(cluster, table)
    got = _engine_run(cluster, table, engine=engine)[0]
    _assert_same(ref, got, name)
    if mut == "huge":
        assert got[5].finished < table.n and got[5].done == 1


def test_empty_trace_is_rejected_like_the_reference():
    """No finished job -> the reference asserts in LogManager.jcts (log_manager.py:138)."""
    from gpuschedule_b200 import capi, ingest, tracegen
    table = ingest.table_from_columns(tracegen.synth_columns(3, seed=1))
    table.n = 0
    for f in ("arrive_tick", "gpus", "gpu_per_task", "duration", "mem_bytes"):
        setattr(table, f, getattr(table, f)[:0])
    with capi.Engine(device=0, nsims=1) as eng:
        eng.config(0, capi.make_cluster(1, 2, 8))
        eng.load_trace(0, table)
        eng.run(0, 0)
        st = eng.stats(0)
        assert st.done == 1 and st.ticks == 0 and st.finished == 0


def test_batched_sweep_writes_reference_bytes(tmp_path):
    """Several golden cases as replicas of ONE launch; each replica's files equal the reference's."""
    import os
    from conftest import GOLDEN
    from gpuschedule_b200 import sweep
    cases = [("kat0", dict(num_switch=4, num_node_p_switch=32)), ("n64", dict(num_switch=4, num_node_p_switch=32)),
             ("sat300", dict(num_switch=1, num_node_p_switch=4)), ("gpc2", dict(num_switch=2, num_node_p_switch=8))]
    sets = [sweep.make_flags(trace_file=os.path.join(GOLDEN, c, "trace.csv"), log_path=c, seed=7, **kw) for c, kw in cases]
    res = sweep.run_batched(sets, out_root=str(tmp_path))
    for (c, _), (out_dir, st) in zip(cases, res):
        for name in ("job.csv", "cluster.csv"):
            got = open(os.path.join(out_dir, name), newline="").read()
            exp = open(os.path.join(GOLDEN, c, name), newline="").read()
            assert got == exp, (c, name)


def test_span_budget_overflow_is_reported_not_written():
    """A span pool sized below the need makes the replica stop with GS_ERR_CAPACITY (no OOB write)."""
    from gpuschedule_b200 import capi, ingest, tracegen
    cluster = capi.make_cluster(num_switch=1, num_node_p_switch=16)
    table = ingest.table_from_columns(tracegen.synth_columns(6000, seed=91, rate=0.3,
                                                             gpu_choices=[16, 32, 64], gpu_probs=[.4, .4, .2]))
    base = _engine_run(cluster, table)[0]                  # default pool: worst case, cannot overflow
    assert len(base[4]) > 60 + 4096                        # more spans than the tiny budget below
    for engine in ENGINES:
        with capi.Engine(device=0, nsims=1) as eng:
            eng.set_engine(engine)
            eng.set_span_budget(0.01)                      # 60 + 4096 records
            eng.config(0, cluster)
            eng.load_trace(0, table)
            with pytest.raises(capi.GsError, match="in-kernel error"):
                eng.run_all()
            assert eng.stats(0).status == -4
        with capi.Engine(device=0, nsims=1) as eng:        # generous budget == default result
            eng.set_engine(engine)
            eng.set_span_budget(8.0)
            eng.config(0, cluster)
            eng.load_trace(0, table)
            rows = eng.run_all()[0]
            assert eng.stats(0).done == 1 and rows.tobytes() == base[0].tobytes()


@pytest.mark.parametrize("engine", ENGINES)
def test_engine_fuzz_matches_oracle(engine):
    """Differential fuzz: 100 random clusters / traces (the generator of tests/test_cpu_differential.py:
    1..64 GPUs per node, 1..132 nodes, cpu- or memory-bound nodes, leaks, gpu_per_container up to 4,
    saturating rates) run as heterogeneous replicas of ONE handle and compared with the oracle."""
    import oracle
    from test_cpu_differential import _case
    from gpuschedule_b200 import capi
    cases = [_case(seed) for seed in range(300, 400)]
    with capi.Engine(device=0, nsims=len(cases)) as eng:
        eng.set_engine(engine)
        for i, (cluster, table) in enumerate(cases):
            eng.config(i, cluster)
            eng.load_trace(i, table)
        rows = eng.run_all()
        for i, (cluster, table) in enumerate(cases):
            ref = oracle.run_fifo(cluster, table)
            recs, order = eng.fetch_jobs(i)
            span_off, spans = eng.fetch_spans(i)
            _assert_same(ref, (rows[i], recs, order, span_off, spans, eng.stats(i)), f"fuzz case {300 + i}")


def test_compact_records_decode_to_the_same_rows():
    """gs_fetch_compact (what the bench's end-to-end path and the CLI read) decoded on the host == the rows the
    device expands for gs_fetch_rows == the oracle; also with a window too small for the run and with a separate,
    tiny capacity for the queue records."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    from gpuschedule_b200 import log_manager as lm
    for seed, rate, ckw in [(31, 0.5, dict(num_switch=4, num_node_p_switch=32)), (32, 2.5, dict(num_switch=1, num_node_p_switch=6)),
                            (33, 1.0, dict(num_switch=16, num_node_p_switch=64)),
                            (34, 1.5, dict(num_switch=1, num_node_p_switch=3, num_gpu_p_node=48, num_cpu_p_node=800, mem_p_node=4000))]:   # > 32 GPUs per node: 16-byte spans
        cluster = capi.make_cluster(**ckw)
        m, g = cluster.num_switch * cluster.num_node_p_switch, cluster.num_gpu_p_node
        table = ingest.table_from_columns(tracegen.synth_columns(2500, seed=seed, rate=rate))
        ref = oracle.run_fifo(cluster, table)
        for cap, qcap in ((0, 0), (50, 3)):
            with capi.Engine(device=0, nsims=1) as eng:
                if qcap:
                    eng.set_queue_rows_cap(qcap)
                eng.config(0, cluster)
                eng.load_trace(0, table)
                parts_dev, parts_host = [], []
                while True:
                    eng.run(0, cap)
                    w, ev, qr, ne, jobs, dur, order, pool = eng.fetch_compact(0)
                    assert len(ev) >= 1 and len(ne) >= 1 and (cap == 0 or (len(ev) <= cap and len(qr) <= qcap))
                    parts_host.append(lm.expand_rows(ev, qr, ne, w.row_first, w.ticks, m, g))
                    parts_dev.append(eng.fetch_rows(0, w.row_first, w.ticks - w.row_first))
                    if eng.stats(0).done:
                        break
                rows_host, rows_dev = np.concatenate(parts_host), np.concatenate(parts_dev)
                assert rows_host.tobytes() == rows_dev.tobytes() == ref.rows.tobytes(), (seed, cap)
                recs = lm.expand_jobs(jobs, int(w.admitted), table.duration)
                assert recs.tobytes() == ref.recs.tobytes() == eng.fetch_jobs(0)[0].tobytes(), (seed, cap)
                assert np.array_equal(order, ref.finish_order)
                off, spans = lm.group_spans(jobs, int(w.admitted), pool)
                assert np.array_equal(off, ref.span_off) and spans.tobytes() == ref.spans.tobytes(), (seed, cap)
                if cap == 0 and rate < 1.0:
                    assert len(ev) < 0.8 * w.ticks          # the jumped ticks left no record


def test_async_load_and_fetch_from_pinned_buffers():
    """gs_set_async: uploads from page-locked buffers are enqueued without staging, gs_fetch_compact enqueues the
    copies of many replicas back to back and one gs_sync waits -- same bytes as the synchronous path."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    from gpuschedule_b200 import log_manager as lm
    from gpuschedule_b200.log_manager import CSPAN_DTYPE, EVROW_DTYPE, JOBRUN_DTYPE, NODEEV_DTYPE, QROW_DTYPE
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=12)
    tables = [ingest.table_from_columns(tracegen.synth_columns(900 + 10 * i, seed=60 + i, rate=0.8)) for i in range(12)]
    refs = [oracle.run_fifo(cluster, t) for t in tables]
    nmax = max(t.n for t in tables)
    pin_in = [capi.PinnedBuffer(nmax * 32) for _ in tables]
    with capi.Engine(device=0, nsims=len(tables)) as eng:
        eng.set_async(True)
        for step in range(2):                              # second step: buffers and slabs are reused
            for i, t in enumerate(tables):
                if step == 0:
                    eng.config(i, cluster)
                v = pin_in[i].view(capi.JOBIN_DTYPE, t.n)
                v[:] = t.packed()
                eng.load_trace_packed(i, v)
            eng.run(0, 0)
            wins = [eng.window(i) for i in range(len(tables))]
            outs = []
            for i, w in enumerate(wins):
                assert eng.result_layout(i).span_bytes == 8          # 8 GPUs per node: the 8-byte span records
                pb = capi.PinnedBuffer(24 * (w.ev_rows + w.q_rows) + 8 * w.node_events + 4 * w.n + 4 * w.finished + 8 * w.spans_used + 64)
                o = 0
                ev = pb.view(EVROW_DTYPE, w.ev_rows, o); o += 24 * w.ev_rows
                qr = pb.view(QROW_DTYPE, w.q_rows, o); o += 24 * w.q_rows
                ne = pb.view(NODEEV_DTYPE, w.node_events, o); o += 8 * w.node_events
                jobs = pb.view(JOBRUN_DTYPE, w.n, o); o += 4 * w.n
                sp = pb.view(CSPAN_DTYPE, w.spans_used, o); o += 8 * w.spans_used
                order = pb.view(np.int32, w.finished, o)
                eng.fetch_compact_into(i, ev, qr, ne, jobs, None, order, sp)
                outs.append((pb, ev, qr, ne, jobs, order, sp))
            eng.sync()
            for i, (w, (pb, ev, qr, ne, jobs, order, sp)) in enumerate(zip(wins, outs)):
                rows = lm.expand_rows(ev, qr, ne, w.row_first, w.ticks, 24, 8)
                assert rows.tobytes() == refs[i].rows.tobytes(), (step, i)
                assert np.array_equal(order, refs[i].finish_order)
                assert lm.expand_jobs(jobs, int(w.admitted), tables[i].duration).tobytes() == refs[i].recs.tobytes()
                off, spans = lm.group_spans(jobs, int(w.admitted), sp)
                assert spans.tobytes() == refs[i].spans.tobytes()
                pb.free()
    for pb in pin_in:
        pb.free()


def test_per_call_registries_step_like_the_fused_loop():
    """scheduling_algorithms['fifo'] / placement_algorithms['yarn'] driven from the host one tick at a time, the way the
    reference's Scheduler.start drives its registries (schedule.py:185-209): same start tick, same nodes and same finish
    order as the fused device loop -- which is pinned to the reference's bytes."""
    import math
    import os
    from types import SimpleNamespace
    from conftest import GOLDEN
    from gpuschedule_b200 import algorithm, capi
    from gpuschedule_b200.infrastructure import Infrastructure
    from gpuschedule_b200.jobs import JobQueueManager, JobsManager
    for case in ("kat0", "n64"):
        table, cluster, meta, _, _ = load_golden(case)
        rows, recs, order, span_off, spans, st = _engine_run(cluster, table)[0]
        flags = SimpleNamespace(num_switch=cluster.num_switch, num_node_p_switch=cluster.num_node_p_switch,
                                num_gpu_p_node=cluster.num_gpu_p_node, num_cpu_p_node=cluster.num_cpu_p_node, mem_p_node=cluster.mem_p_node,
                                gpu_memory_capacity=cluster.gpu_mem_cap_mib // 1024, enable_network_costs=False, bandwidth=1250,
                                internode_latency=0.015, schedule="fifo", scheme="yarn", num_queue=1,
                                trace_file=os.path.join(GOLDEN, case, "trace.csv"))
        infra = Infrastructure(flags)
        jm = JobsManager(flags, JobQueueManager(flags))
        running, finished, delta = [], [], 0
        while jm.remaining_jobs() + len(running) > 0:
            jm.gen_jobs(delta)
            if jm.queuing_jobs() > 0:
                nodes, job, ok = algorithm.scheduling_algorithms["fifo"]("yarn", algorithm.placement_algorithms["yarn"], infra, jm, delta)
                if ok:
                    job.start_time = delta
                    job.end_time = delta + max(1, math.ceil(job.duration))
                    running.append(job)
                    assert sorted(int(k) - 1 for k in nodes) == sorted(spans["node"][span_off[job.index]:span_off[job.index + 1]].tolist())
            delta += 1
            for job in [j for j in running if j.end_time == delta]:
                algorithm.release_job(infra, job)
                running.remove(job)
                finished.append(job)
        assert delta == st.ticks
        assert [j.index for j in finished] == order.tolist()
        for j in finished:
            assert (j.start_time, j.end_time) == (int(recs["start"][j.index]), int(recs["end"][j.index]))


def test_batch_upload_and_strided_read_back():
    """gs_load_traces_packed (all traces of a handle, one strided upload) + gs_fetch_results (all result blocks, one strided
    copy) -- the path bench.py's end-to-end measurement uses -- against the oracle, replica by replica, with traces of
    different lengths and a second step that reuses the arenas."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    from gpuschedule_b200 import log_manager as lm
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=12)
    tables = [ingest.table_from_columns(tracegen.synth_columns(700 + 37 * i, seed=90 + i, rate=0.6 + 0.2 * (i % 4))) for i in range(9)]
    refs = [oracle.run_fifo(cluster, t) for t in tables]
    nmax = max(t.n for t in tables)
    pin = capi.PinnedBuffer(len(tables) * nmax * 32)
    block = pin.view(capi.JOBIN_DTYPE, len(tables) * nmax)
    for i, t in enumerate(tables):
        block[i * nmax:i * nmax + t.n] = t.packed()
    for use_async in (False, True):
        with capi.Engine(device=0, nsims=len(tables)) as eng:
            eng.set_async(use_async)
            for i in range(len(tables)):
                eng.config(i, cluster)
            for step in range(2):
                eng.load_traces_packed(block, nmax * 32, [t.n for t in tables])
                eng.run(0, 0)
                lay = [eng.result_layout(i) for i in range(len(tables))]
                pitch = max(int(x.block_bytes) for x in lay)
                out = np.zeros(len(tables) * pitch, dtype=np.uint8)
                eng.fetch_results(out, pitch)
                eng.sync()
                for i, t in enumerate(tables):
                    w = eng.window(i)
                    ev, qr, ne, jobs, order, pool = capi.Engine.result_views(out, pitch, i, lay[i], w)
                    rows = lm.expand_rows(ev, qr, ne, w.row_first, w.ticks, 24, 8)
                    assert eng.stats(i).done == 1 and rows.tobytes() == refs[i].rows.tobytes(), (use_async, step, i)
                    assert lm.expand_jobs(jobs, int(w.admitted), t.duration).tobytes() == refs[i].recs.tobytes()
                    assert np.array_equal(order, refs[i].finish_order)
                    off, spans = lm.group_spans(jobs, int(w.admitted), pool)
                    assert np.array_equal(off, refs[i].span_off) and spans.tobytes() == refs[i].spans.tobytes()
                    recs2, order2 = eng.fetch_jobs(i)               # the per-replica calls see the same arena
                    assert recs2.tobytes() == refs[i].recs.tobytes() and np.array_equal(order2, order)
    pin.free()
