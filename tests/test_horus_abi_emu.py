"""The utilisation-aware engine's C ABI on a CPU-only box.

tests/emu builds gs_horus.cu ITSELF with g++ against a stand-in cuda_runtime.h (device memory = host memory, the
kernel launch = the source's own host loop over the same device functions).  Everything the GPU tests do through
capi.HorusEngine is therefore repeated here against the real entry points -- gs_horus_create / config / load_trace /
load_stream / load_words / set_lanes / run / stats / fetch: the slab layout of prepare(), the stream and word-table
uploads, resumed launches, per-replica restarts after GS_ERR_CAPACITY, the error codes -- with the reference
fixtures and the oracle as judges.  What remains GPU-only is the device compilation of the same functions and the
launch itself (tests/test_gpu_widen_horus.py).  This is test infrastructure: the package never loads this library."""
import functools
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN, REPO, horus_cases, load_horus, render_horus_outputs


@pytest.fixture(scope="module")
def engine_cls():
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("tests_emu", os.path.join(REPO, "tests", "emu", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["tests_emu"] = mod
    spec.loader.exec_module(mod)
    return mod.emu_engine_class()


@functools.lru_cache(maxsize=4)
def _stream(seed, count=1 << 21):
    np.random.seed(seed)
    return np.random.standard_normal(count)


@functools.lru_cache(maxsize=4)
def _words(seed, count=6 << 20):
    np.random.seed(seed)
    return np.random.randint(0, 2 ** 32, size=count, dtype=np.uint32)


def _collect(eng, i):
    rows, util, flags, recs, order = eng.fetch(i)
    st = eng.stats(i)
    return SimpleNamespace(rows=rows, util=util, util_is_array=flags, recs=recs, finish_order=order,
                           events=int(st.events), draws=int(st.draws), ticks=int(st.ticks), done=int(st.done))


def _run(engine_cls, jobs, max_ticks=0, lanes=1, words=False):
    from gpuschedule_b200 import capi
    with engine_cls(device=0, nsims=len(jobs)) as eng:
        eng.set_lanes(lanes)
        for i, (cluster, table, params) in enumerate(jobs):
            eng.config(i, cluster, capi.make_horus_params(params["scheme"], params["schedule"], params["num_buffer"], params.get("num_queue", 1)))
            eng.load_trace(i, table)
            if words or params["schedule"] == "horus+":
                eng.load_words(i, _words(params["seed"]))
            else:
                eng.load_stream(i, _stream(params["seed"]))
        launches = 0
        while True:
            eng.run(max_ticks=max_ticks, rows_cap=1 << 15)
            launches += 1
            if all(eng.stats(i).done for i in range(len(jobs))):
                break
            assert launches < 10000
        return [_collect(eng, i) for i in range(len(jobs))], launches


@pytest.mark.parametrize("lanes,words,max_ticks", [(1, False, 0), (32, False, 50), (0, False, 0), (0, True, 64), (1, True, 0)],
                         ids=["scalar", "scalar-resumed", "cooperative", "cooperative-words-resumed", "scalar-words"])
def test_every_fixture_through_the_c_abi(engine_cls, lanes, words, max_ticks):
    cases = horus_cases()
    loaded = [load_horus(c) for c in cases]
    results, launches = _run(engine_cls, [(cl, tb, pr) for tb, cl, pr, _, _ in loaded], max_ticks=max_ticks, lanes=lanes, words=words)
    assert (launches > 3) == (max_ticks > 0)
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv, case
        assert got_cluster == cluster_csv, case


def test_gpu_test_scenarios_match_the_oracle(engine_cls):
    """the 40 heterogeneous replicas of tests/test_gpu_widen_horus.py, one handle, every kernel mapping"""
    import oracle
    import test_gpu_widen_horus as gpu_tests
    jobs = [gpu_tests._seeded_case(s) for s in range(40)]
    refs = [oracle.run_horus(cluster, table, **params) for cluster, table, params in jobs]
    for lanes in (1, 32, 0):
        results, _ = _run(engine_cls, jobs, lanes=lanes)
        for s, (ref, res) in enumerate(zip(refs, results)):
            assert res.ticks == ref.ticks and res.draws == ref.draws and res.events == ref.events, (lanes, s)
            assert res.rows.tobytes() == ref.rows.tobytes() and res.util.tobytes() == ref.util.tobytes(), (lanes, s)
            assert res.util_is_array.tobytes() == ref.util_is_array.tobytes() and res.recs.tobytes() == ref.recs.tobytes(), (lanes, s)
            assert np.array_equal(res.finish_order, ref.finish_order), (lanes, s)


def test_shared_streams_and_error_codes(engine_cls):
    import oracle
    from gpuschedule_b200 import capi
    table, cluster, params, _, _ = load_horus("horus_small")
    t2, c2, p2, _, _ = load_horus("horus_buf1")
    with engine_cls(device=0, nsims=2) as eng:
        eng.config(0, cluster, capi.make_horus_params("horus", "horus", params["num_buffer"]))
        eng.config(1, c2, capi.make_horus_params("horus", "horus", p2["num_buffer"]))
        eng.load_trace(0, table)
        eng.load_trace(1, t2)
        eng.load_stream(-1, _stream(7, 1000))                          # one stream for both, too short
        with pytest.raises(capi.GsError) as e:
            eng.run(rows_cap=1 << 15)
        assert e.value.code == capi.GS_ERR_CAPACITY and eng.stats(0).status == capi.GS_ERR_CAPACITY
        eng.load_stream(-1, _stream(7))                                # both replicas read the same seed-7 stream from 0
        eng.run(rows_cap=1 << 15)
        for i, (tb, cl, pr) in enumerate(((table, cluster, params), (t2, c2, p2))):
            ref = oracle.run_horus(cl, tb, **pr)
            res = _collect(eng, i)
            assert res.done == 1 and res.rows.tobytes() == ref.rows.tobytes() and res.util.tobytes() == ref.util.tobytes()
        eng.load_words(-1, _words(7))                                  # the same stream as raw words: same answer
        eng.run(rows_cap=1 << 15)
        ref = oracle.run_horus(cluster, table, **params)
        assert _collect(eng, 0).rows.tobytes() == ref.rows.tobytes()
        with pytest.raises(capi.GsError) as e:                          # rows_cap too small
            eng.load_stream(-1, _stream(7))
            eng.run(rows_cap=10)
        assert e.value.code == capi.GS_ERR_CAPACITY and eng.stats(0).ticks == 10
        with pytest.raises(capi.GsError):
            eng.config(0, cluster, capi.GsHorusParams(0, 7, 5, 0))      # unknown schedule
        with pytest.raises(capi.GsError):
            eng.config(0, cluster, capi.GsHorusParams(1, 1, 5, 0))      # gandiva_score under schedule horus: the reference cannot do that
        eng.config(0, cluster, capi.make_horus_params("horus+", "horus+", 15, 3))
        eng.load_stream(0, _stream(7))
        with pytest.raises(capi.GsError) as e:                          # horus+ needs the raw word stream
            eng.run(rows_cap=1 << 15)
        assert e.value.code == capi.GS_ERR_STATE


def test_host_mirror_and_batched_sweep_through_the_c_abi(engine_cls, tmp_path, monkeypatch):
    """Scheduler.start() and sweep.run_batched_horus end to end on the real ABI (small chunks: restarts of single replicas)"""
    from gpuschedule_b200 import capi, infrastructure, jobs, log_manager, schedule, sweep
    monkeypatch.setattr(capi, "HorusEngine", engine_cls)
    cases = ["horus_small", "gandiva_slice", "horusplus_k3", "yarn_sched_horus"]
    sets = []
    for case in cases:
        table, cluster, params, _, _ = load_horus(case)
        sets.append(sweep.make_flags(trace_file=os.path.join(GOLDEN, case, "trace.csv"), scheme=params["scheme"], schedule=params["schedule"],
                                     num_buffer=params["num_buffer"], num_queue=params["num_queue"], num_switch=cluster.num_switch,
                                     num_node_p_switch=cluster.num_node_p_switch, num_gpu_p_node=cluster.num_gpu_p_node,
                                     seed=params["seed"], log_path=case))
    results = sweep.run_batched_horus(sets, out_root=str(tmp_path / "sweep"), chunk=150000)
    for case, (out_dir, st) in zip(cases, results):
        _, _, _, job_csv, cluster_csv = load_horus(case)
        assert open(os.path.join(out_dir, "job.csv"), newline="").read() == job_csv, case
        assert open(os.path.join(out_dir, "cluster.csv"), newline="").read() == cluster_csv, case
    fl = sets[0]
    infra = infrastructure.Infrastructure(fl)
    jm = jobs.JobsManager(fl, jobs.JobQueueManager(fl, fl.trace_file))
    os.makedirs(tmp_path / "single")
    lm = log_manager.LogManager(str(tmp_path / "single"), fl)
    lm.init(infra)
    np.random.seed(fl.seed)
    schedule.Scheduler(infra, jm, lm).start()
    _, _, _, job_csv, cluster_csv = load_horus(cases[0])
    assert open(tmp_path / "single" / "job.csv", newline="").read() == job_csv
    assert open(tmp_path / "single" / "cluster.csv", newline="").read() == cluster_csv


def test_scheme_schedule_matrix_of_make_horus_params():
    """the same matrix tests/test_gpu_widen_horus.py asserts on the device (algorithm.py:182-187,292-298): yarn is a served
    placement under the look-ahead schedulers, fifo has no score function (KeyError at algorithm.py:58), unknown names die"""
    from gpuschedule_b200 import capi
    yp = capi.make_horus_params("yarn", "horus", 5)
    assert (yp.placement, yp.schedule, yp.score) == (1, 1, 0)
    assert capi.make_horus_params("horus", "gandiva", 5).placement == 0
    assert capi.make_horus_params("gandiva", "gandiva", 5).score == 1
    for scheme, schedule in (("yarn", "fifo"), ("random", "horus"), ("horus", "sjf")):
        with pytest.raises(NotImplementedError):
            capi.make_horus_params(scheme, schedule, 5)
