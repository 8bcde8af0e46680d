"""CUDA engine for the utilisation-aware paths (include/gsched_horus.h) vs the reference fixtures
(tests/golden/horus_* / gandiva_* / horusplus_*, bytes of job.csv and cluster.csv incl. the sampled utilisation
column) and vs the pinned oracle (oracle/horus_oracle.c) on seeded cases; several replicas per launch, resumed
runs, both stream forms, and the too-short-stream error.  (The file name sorts after the main-path GPU tests on
purpose: the widening row must never keep `pytest -x` from reaching them.)

Every test here is a hard assertion: nothing is marked xfail.  The same device functions also run on the CPU
through tests/emu (tests/test_horus_emu.py, tests/test_horus_abi_emu.py)."""
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import horus_cases, load_horus, render_horus_outputs

pytestmark = pytest.mark.gpu


def _stream(seed, count=1 << 21):
    np.random.seed(seed)
    return np.random.standard_normal(count)


def _words(seed, count=6 << 20):
    np.random.seed(seed)
    return np.random.randint(0, 2 ** 32, size=count, dtype=np.uint32)        # raw MT19937 output words


def _collect(eng, i):
    rows, util, flags, recs, order = eng.fetch(i)
    st = eng.stats(i)
    return SimpleNamespace(rows=rows, util=util, util_is_array=flags, recs=recs, finish_order=order,
                           events=int(st.events), draws=int(st.draws), ticks=int(st.ticks), done=int(st.done))


def _run(jobs, max_ticks=0, lanes=1, words=False):
    """jobs: list of (cluster, table, params); one replica each, one handle."""
    from gpuschedule_b200 import capi
    with capi.HorusEngine(device=0, nsims=len(jobs)) as eng:
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


RAN_GREEN_IN_ROUND_1 = ["gandiva_slice", "gandiva_small", "horus_buf1", "horus_racks", "horus_small"]


def _served(plus=False):
    return [c for c in horus_cases() if (load_horus(c)[2]["schedule"] == "horus+") == plus]


def test_engine_matches_reference_bytes_all_fixtures_one_launch():
    cases = RAN_GREEN_IN_ROUND_1
    loaded = [load_horus(c) for c in cases]
    results, launches = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded], lanes=32)
    assert launches == 1
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv, case
        assert got_cluster == cluster_csv, case


def test_engine_resumed_every_50_ticks_is_identical():
    table, cluster, params, job_csv, cluster_csv = load_horus("gandiva_slice")
    (res,), launches = _run([(cluster, table, params)], max_ticks=50)
    assert launches > 5
    got_job, got_cluster = render_horus_outputs(table, cluster, res)
    assert got_job == job_csv and got_cluster == cluster_csv


def _seeded_case(seed):
    from gpuschedule_b200 import capi, ingest, tracegen
    rng = np.random.default_rng(300 + seed)
    kind = ["horus", "gandiva"][seed % 2]
    G = int(rng.choice([2, 4, 8]))
    gpc = 2 if seed % 5 == 4 else 1
    cluster = capi.make_cluster(num_switch=int(rng.integers(1, 4)), num_node_p_switch=int(rng.integers(1, 5)), num_gpu_p_node=G,
                                num_cpu_p_node=int(rng.choice([36, 60, 128])), mem_p_node=int(rng.choice([180, 300, 512])),
                                gpu_memory_capacity=int(rng.choice([16, 32])))
    choices = sorted(set(int(x) * gpc for x in rng.choice([1, 1, 2, 3, 4, 6, 8], size=4)))
    table = ingest.table_from_columns(tracegen.synth_columns(int(rng.integers(20, 150)), seed=700 + seed, rate=float(rng.choice([1.0, 2.0, 4.0])),
                                                             gpu_per_container=gpc, gpu_choices=choices, gpu_probs=rng.dirichlet(np.ones(len(choices))),
                                                             max_mem_mib=int(rng.choice([8000, 16384, 33500]))))
    return cluster, table, dict(scheme=kind, schedule=kind, num_buffer=int(rng.choice([1, 3, 5])), num_queue=1, seed=2000 + seed)


@pytest.mark.parametrize("lanes", [1, 32])
def test_engine_matches_oracle_on_40_heterogeneous_replicas(lanes):
    import oracle
    jobs = [_seeded_case(s) for s in range(40)]
    results, _ = _run(jobs, lanes=lanes)
    for s, ((cluster, table, params), res) in enumerate(zip(jobs, results)):
        ref = oracle.run_horus(cluster, table, **params)
        assert res.ticks == ref.ticks and res.draws == ref.draws and res.events == ref.events, s
        assert res.rows.tobytes() == ref.rows.tobytes(), s
        assert res.util.tobytes() == ref.util.tobytes() and res.util_is_array.tobytes() == ref.util_is_array.tobytes(), s
        assert res.recs.tobytes() == ref.recs.tobytes(), s
        assert np.array_equal(res.finish_order, ref.finish_order), s


def test_short_stream_and_unserved_schedule_fail_loudly():
    from gpuschedule_b200 import capi
    table, cluster, params, _, _ = load_horus("horus_small")
    with capi.HorusEngine(device=0, nsims=1) as eng:
        eng.config(0, cluster, capi.make_horus_params("horus", "horus", 5))
        eng.load_trace(0, table)
        eng.load_stream(0, _stream(params["seed"], 1000))
        with pytest.raises(capi.GsError) as e:
            eng.run(rows_cap=1 << 15)
        assert e.value.code == -4
        eng.load_stream(0, _stream(params["seed"]))          # a longer stream: the run starts over and completes
        eng.run(rows_cap=1 << 15)
        assert eng.stats(0).done == 1
        with pytest.raises(capi.GsError):
            eng.config(0, cluster, capi.GsHorusParams(0, 7, 5, 0))       # unknown schedule
        with pytest.raises(capi.GsError):
            eng.config(0, cluster, capi.GsHorusParams(0, 2, 5, 0))       # horus+ without queues
        eng.config(0, cluster, capi.make_horus_params("horus+", "horus+", 15, 3))
        with pytest.raises(capi.GsError) as e:                           # horus+ needs the raw word stream
            eng.run(rows_cap=1 << 15)
        assert e.value.code == -3
    # scheme x schedule matrix of the utilisation-aware engine (algorithm.py:182-187,292-298): yarn is served
    # (ms_yarn_placement under the look-ahead schedulers); fifo is not (KeyError in score_fn[schedule], algorithm.py:58)
    yp = capi.make_horus_params("yarn", "horus", 5)
    assert (yp.placement, yp.schedule, yp.score) == (1, 1, 0)
    assert capi.make_horus_params("horus", "gandiva", 5).placement == 0
    with pytest.raises(NotImplementedError):
        capi.make_horus_params("yarn", "fifo", 5)
    with pytest.raises(NotImplementedError):
        capi.make_horus_params("random", "horus", 5)


@pytest.mark.parametrize("case", ["horus_racks", "gandiva_small"])
def test_cli_run_sim_horus_writes_reference_bytes(case, tmp_path):
    """run_sim.py --scheme horus|gandiva end to end: same flags as the reference, same files, same bytes."""
    import glob
    import json
    import os
    import shutil
    import subprocess
    import sys
    from conftest import GOLDEN, REPO
    f = json.load(open(os.path.join(GOLDEN, case, "horus.json")))
    fl = f["flags"]
    shutil.copy(os.path.join(GOLDEN, case, "trace.csv"), tmp_path / "trace.csv")
    cmd = [sys.executable, os.path.join(REPO, "run_sim.py"), "--scheme", fl["_scheme"], "--schedule", fl["_schedule"],
           "--trace_file", "trace.csv", "--log_path", "g", "--seed", str(f["numpy_seed"])]
    for k, v in fl.items():
        if not k.startswith("_"):
            cmd += ["--" + k, str(v)]
    subprocess.run(cmd, cwd=tmp_path, check=True, capture_output=True)
    runs = glob.glob(str(tmp_path / "log" / "g" / "*"))
    assert len(runs) == 1
    for name in ("job.csv", "cluster.csv"):
        got = open(os.path.join(runs[0], name), newline="").read()
        exp = open(os.path.join(GOLDEN, case, name), newline="").read()
        assert got == exp, name


def test_horus_plus_matches_reference_bytes():
    cases = _served(plus=True)
    assert len(cases) >= 2
    loaded = [load_horus(c) for c in cases]
    results, _ = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded])
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv, case
        assert got_cluster == cluster_csv, case


def test_word_stream_serves_horus_and_gandiva_too():
    cases = _served()
    loaded = [load_horus(c) for c in cases]
    results, _ = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded], words=True)
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv and got_cluster == cluster_csv, case


def test_batched_sweep_on_device(tmp_path):
    """sweep.run_batched_horus: seeded horus / gandiva / horus+ replicas in one launch == the single-run fixtures."""
    import os
    from conftest import GOLDEN
    from gpuschedule_b200 import sweep
    cases = ["horus_small", "gandiva_slice", "horusplus_k3"]
    sets = []
    for case in cases:
        table, cluster, params, _, _ = load_horus(case)
        sets.append(sweep.make_flags(trace_file=os.path.join(GOLDEN, case, "trace.csv"), scheme=params["scheme"], schedule=params["schedule"],
                                     num_buffer=params["num_buffer"], num_queue=params["num_queue"], num_switch=cluster.num_switch,
                                     num_node_p_switch=cluster.num_node_p_switch, num_gpu_p_node=cluster.num_gpu_p_node,
                                     seed=params["seed"], log_path=case))
    results = sweep.run_batched_horus(sets, out_root=str(tmp_path), chunk=200000)      # small chunks: single replicas restart
    for case, (out_dir, st) in zip(cases, results):
        _, _, _, job_csv, cluster_csv = load_horus(case)
        assert open(os.path.join(out_dir, "job.csv"), newline="").read() == job_csv, case
        assert open(os.path.join(out_dir, "cluster.csv"), newline="").read() == cluster_csv, case


def test_crossed_and_yarn_fixtures_match_reference_bytes():
    """score function by schedule name (cross_*) and --scheme yarn under these schedulers (yarn_sched_*)"""
    cases = [c for c in _served() if c not in RAN_GREEN_IN_ROUND_1]
    assert len(cases) >= 4
    loaded = [load_horus(c) for c in cases]
    results, _ = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded])
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv and got_cluster == cluster_csv, case


def test_cooperative_warp_mapping_matches_oracle_and_fixtures():
    """gs_horus_set_lanes(0): one simulation per warp, all lanes score a candidate job's devices together."""
    import oracle
    jobs = [_seeded_case(s) for s in range(24)]
    results, _ = _run(jobs, lanes=0)
    for s, ((cluster, table, params), res) in enumerate(zip(jobs, results)):
        ref = oracle.run_horus(cluster, table, **params)
        assert res.ticks == ref.ticks and res.draws == ref.draws and res.events == ref.events, s
        assert res.rows.tobytes() == ref.rows.tobytes() and res.util.tobytes() == ref.util.tobytes(), s
        assert res.recs.tobytes() == ref.recs.tobytes() and np.array_equal(res.finish_order, ref.finish_order), s
    loaded = [load_horus(c) for c in RAN_GREEN_IN_ROUND_1]
    results, _ = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded], lanes=0, max_ticks=64)      # resumed every 64 ticks
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(RAN_GREEN_IN_ROUND_1, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv and got_cluster == cluster_csv, case


def test_cooperative_warp_mapping_horus_plus():
    cases = _served(plus=True)
    loaded = [load_horus(c) for c in cases]
    results, _ = _run([(cl, tb, pr) for tb, cl, pr, _, _ in loaded], lanes=0)
    for case, (table, cluster, params, job_csv, cluster_csv), res in zip(cases, loaded, results):
        got_job, got_cluster = render_horus_outputs(table, cluster, res)
        assert got_job == job_csv and got_cluster == cluster_csv, case
