"""CPU check of the logic the horus / gandiva CUDA kernel runs.

gpuschedule_b200/csrc/gs_horus_core.cuh is scalar __host__ __device__ code; tests/emu/ compiles that very
header with g++ and this file compares it, fed with numpy's own standard-normal stream, against
  * the reference fixtures (tests/golden/horus_* / gandiva_*: job.csv and all 13 cluster.csv columns), and
  * the pinned oracle (oracle/horus_oracle.c, which carries its own MT19937) on seeded random cases.
It also proves the stream contract of include/gsched_horus.h: numpy.random.normal(loc, scale, size=1)
number k equals loc + scale * standard_normal()[k].  The device build of the same functions is checked
on the GPU by tests/test_gpu_widen_horus.py."""
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import horus_cases, load_horus, render_horus_outputs


def _stream(seed, count):
    np.random.seed(seed)
    return np.random.standard_normal(count)


def _words(seed, count):
    np.random.seed(seed)
    return np.random.randint(0, 2 ** 32, size=count, dtype=np.uint32)       # the raw MT19937 output words


def _emu(table, cluster, params, count=1 << 21, step=0, use_words=None, cooperative=False):
    from gpuschedule_b200 import capi
    from tests_emu import run_horus
    hp = capi.make_horus_params(params["scheme"], params["schedule"], params["num_buffer"], params.get("num_queue", 1))
    if use_words is None:
        use_words = params["schedule"] == "horus+"
    if use_words:
        out = run_horus(cluster, hp, table, None, 1 << 15, step, words=_words(params["seed"], 3 * count), cooperative=cooperative)
    else:
        out = run_horus(cluster, hp, table, _stream(params["seed"], count), 1 << 15, step, cooperative=cooperative)
    ticks, rows, util, flags, recs, order, events, draws = out
    assert ticks >= 0, ticks
    return SimpleNamespace(rows=rows, util=util, util_is_array=flags, recs=recs, finish_order=order, events=events, draws=draws, ticks=ticks)


@pytest.fixture(scope="module", autouse=True)
def _emu_module():
    import importlib.util
    import os
    import sys
    from conftest import REPO
    spec = importlib.util.spec_from_file_location("tests_emu", os.path.join(REPO, "tests", "emu", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["tests_emu"] = mod
    spec.loader.exec_module(mod)
    mod.build()
    yield


def test_numpy_stream_contract():
    np.random.seed(11)
    g = np.random.standard_normal(2000)
    np.random.seed(11)
    rng = np.random.default_rng(5)
    for k in range(2000):
        loc, scale = float(rng.uniform(0, 100)), float(rng.uniform(0, 20))
        assert np.random.normal(loc=loc, scale=scale, size=1)[0] == loc + scale * g[k]


@pytest.mark.parametrize("stream", ["values", "words", "values-cooperative", "words-cooperative"])
@pytest.mark.parametrize("case", [c for c in horus_cases()])
def test_kernel_logic_matches_reference_bytes(case, stream):
    """values / words: the scalar driver with either stream form; values-cooperative: the warp-cooperative driver
    (scoring by sample index, lane loop run sequentially on the host)."""
    table, cluster, params, job_csv, cluster_csv = load_horus(case)
    if params["schedule"] == "horus+" and not stream.startswith("words"):
        pytest.skip("horus+ draws integers too: it needs the raw word stream")
    res = _emu(table, cluster, params, use_words=stream.startswith("words"), cooperative=stream.endswith("cooperative"))
    got_job, got_cluster = render_horus_outputs(table, cluster, res)
    assert got_job == job_csv
    assert got_cluster == cluster_csv


@pytest.mark.parametrize("seed", range(6))
def test_horus_plus_logic_matches_oracle_seeded(seed):
    """k-means re-clustering, credit queues and the integer draws; seed 5 queues > 128 jobs per cluster (numpy's
    pairwise summation takes its recursive branch there)."""
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    rng = np.random.default_rng(900 + seed)
    big = seed == 5
    cluster = capi.make_cluster(num_switch=int(rng.integers(1, 3)), num_node_p_switch=1 if big else int(rng.integers(1, 5)),
                                num_gpu_p_node=int(rng.choice([4, 8])), gpu_memory_capacity=int(rng.choice([16, 32])))
    n = 420 if big else int(rng.integers(30, 140))
    table = ingest.table_from_columns(tracegen.synth_columns(n, seed=800 + seed, rate=8.0 if big else float(rng.choice([1.0, 3.0])),
                                                             gpu_choices=[1, 2, 4], gpu_probs=[.5, .3, .2]))
    params = dict(scheme="horus+", schedule="horus+", num_buffer=int(rng.choice([2, 5, 15])), num_queue=int(rng.integers(2, 6)), seed=3000 + seed)
    ref = oracle.run_horus(cluster, table, **params)
    res = _emu(table, cluster, params, count=1 << 22)
    if big:
        assert int(ref.rows["queued"].max()) > 300
    assert res.ticks == ref.ticks and res.draws == ref.draws and res.events == ref.events
    assert res.rows.tobytes() == ref.rows.tobytes()
    assert res.util.tobytes() == ref.util.tobytes() and res.util_is_array.tobytes() == ref.util_is_array.tobytes()
    assert res.recs.tobytes() == ref.recs.tobytes()
    assert np.array_equal(res.finish_order, ref.finish_order)


@pytest.mark.parametrize("seed", range(12))
def test_kernel_logic_matches_oracle_seeded(seed):
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    rng = np.random.default_rng(100 + seed)
    kind = ["horus", "gandiva"][seed % 2]
    G = int(rng.choice([2, 4, 8]))
    gpc = 2 if seed % 5 == 4 else 1
    cluster = capi.make_cluster(num_switch=int(rng.integers(1, 4)), num_node_p_switch=int(rng.integers(1, 5)), num_gpu_p_node=G,
                                num_cpu_p_node=int(rng.choice([36, 60, 128])), mem_p_node=int(rng.choice([180, 300, 512])),
                                gpu_memory_capacity=int(rng.choice([16, 32])))
    choices = sorted(set(int(x) * gpc for x in rng.choice([1, 1, 2, 3, 4, 6, 8], size=4)))
    table = ingest.table_from_columns(tracegen.synth_columns(int(rng.integers(20, 120)), seed=500 + seed, rate=float(rng.choice([1.0, 2.0, 4.0])),
                                                             gpu_per_container=gpc, gpu_choices=choices, gpu_probs=rng.dirichlet(np.ones(len(choices))),
                                                             max_mem_mib=int(rng.choice([8000, 16384, 33500]))))
    params = dict(scheme=kind, schedule=kind, num_buffer=int(rng.choice([1, 3, 5])), num_queue=1, seed=1000 + seed)
    ref = oracle.run_horus(cluster, table, **params)
    for step, use_words, coop in ((0, False, False), (37, False, False), (0, True, False), (0, False, True), (37, False, True)):
        res = _emu(table, cluster, params, step=step, use_words=use_words, cooperative=coop)   # scalar / resumed / words / cooperative
        assert res.ticks == ref.ticks and res.draws == ref.draws and res.events == ref.events
        assert res.rows.tobytes() == ref.rows.tobytes()
        assert res.util.tobytes() == ref.util.tobytes() and res.util_is_array.tobytes() == ref.util_is_array.tobytes()
        assert res.recs.tobytes() == ref.recs.tobytes()
        assert np.array_equal(res.finish_order, ref.finish_order)


def test_short_stream_is_reported():
    from gpuschedule_b200 import capi
    from tests_emu import run_horus
    table, cluster, params, _, _ = load_horus("horus_small")
    hp = capi.make_horus_params(params["scheme"], params["schedule"], params["num_buffer"])
    ticks = run_horus(cluster, hp, table, _stream(params["seed"], 1000), 1 << 15)[0]
    assert ticks == -4                                     # GS_ERR_CAPACITY: load a longer stream
    ticks = run_horus(cluster, hp, table, None, 1 << 15, words=_words(params["seed"], 3000))[0]
    assert ticks == -4


class _EmuHorusEngine:
    """Same interface as capi.HorusEngine, backed by the host build of the device functions: lets the host-side
    mirror (Scheduler.start / sweep.run_batched_horus -> stream chunks, retry on GS_ERR_CAPACITY, formatting, file
    writing) run on CPU.  Like the library, a finished replica keeps its results across later run() calls."""

    def __init__(self, device=0, nsims=1):
        self.n = nsims
        self.cfg, self.table, self.g, self.w, self.res, self.dirty = ([None] * nsims for _ in range(6))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass

    def config(self, sim, cluster, params):
        self.cfg[sim] = (cluster, params)

    def load_trace(self, sim, table):
        self.table[sim] = table

    def load_stream(self, sim, g):
        self.g[sim], self.w[sim], self.dirty[sim] = np.array(g), None, True

    def load_words(self, sim, w):
        self.g[sim], self.w[sim], self.dirty[sim] = None, np.array(w), True

    def run(self, max_ticks=0, rows_cap=1 << 16):
        from gpuschedule_b200 import capi
        from tests_emu import run_horus
        worst = 0
        for i in range(self.n):
            if self.dirty[i]:
                self.res[i] = run_horus(self.cfg[i][0], self.cfg[i][1], self.table[i], self.g[i], rows_cap, words=self.w[i])
                self.dirty[i] = False
            if self.res[i][0] < 0 and worst == 0:
                worst = self.res[i][0]
        if worst:
            raise capi.GsError("emulated gs_horus_run failed", worst)

    def stats(self, sim):
        r = self.res[sim]
        return SimpleNamespace(ticks=max(r[0], 0), events=r[6], draws=r[7], finished=len(r[5]), done=int(r[0] >= 0),
                               status=min(r[0], 0), kernel_ms=0.0)

    def fetch(self, sim):
        r = self.res[sim]
        return r[1], r[2], r[3], r[4], r[5]


def test_batched_sweep_writes_reference_bytes_per_replica(tmp_path, monkeypatch):
    """sweep.run_batched: horus, gandiva and horus+ replicas in one (emulated) launch, each with its own seeded
    numpy stream handed over in small chunks (forcing restarts of single replicas), next to a plain fifo replica that
    must be routed to the other engine."""
    import glob
    import os
    from conftest import GOLDEN
    from gpuschedule_b200 import capi, sweep
    cases = ["horus_small", "gandiva_slice", "horusplus_k3"]
    sets = []
    for case in cases:
        table, cluster, params, _, _ = load_horus(case)
        sets.append(sweep.make_flags(trace_file=os.path.join(GOLDEN, case, "trace.csv"), scheme=params["scheme"], schedule=params["schedule"],
                                     num_buffer=params["num_buffer"], num_queue=params["num_queue"], num_switch=cluster.num_switch,
                                     num_node_p_switch=cluster.num_node_p_switch, num_gpu_p_node=cluster.num_gpu_p_node,
                                     seed=params["seed"], log_path=case))
    monkeypatch.setattr(capi, "HorusEngine", _EmuHorusEngine)
    results = sweep.run_batched_horus(sets, out_root=str(tmp_path), chunk=30000)
    assert len(results) == 3
    for case, (out_dir, st) in zip(cases, results):
        _, _, _, job_csv, cluster_csv = load_horus(case)
        assert open(os.path.join(out_dir, "job.csv"), newline="").read() == job_csv, case
        assert open(os.path.join(out_dir, "cluster.csv"), newline="").read() == cluster_csv, case
    assert all(sweep._is_utilisation_aware(fl) for fl in sets) and not sweep._is_utilisation_aware(sweep.make_flags())


@pytest.mark.parametrize("case", ["horus_small", "gandiva_slice", "horusplus_k3"])
def test_host_mirror_writes_reference_bytes(case, tmp_path, monkeypatch):
    """Scheduler.start() for --scheme horus|gandiva: numpy's global stream is handed to the engine in chunks (a
    first chunk that is too short makes the run start over with a longer one) and the files equal the reference's."""
    import os
    import shutil
    from conftest import GOLDEN
    from gpuschedule_b200 import capi, infrastructure, jobs, log_manager, schedule, sweep
    table, cluster, params, job_csv, cluster_csv = load_horus(case)
    shutil.copy(os.path.join(GOLDEN, case, "trace.csv"), tmp_path / "trace.csv")
    fl = sweep.make_flags(trace_file=str(tmp_path / "trace.csv"), scheme=params["scheme"], schedule=params["schedule"],
                          num_buffer=params["num_buffer"], num_queue=params["num_queue"], num_switch=cluster.num_switch,
                          num_node_p_switch=cluster.num_node_p_switch, num_gpu_p_node=cluster.num_gpu_p_node)
    monkeypatch.setattr(capi, "HorusEngine", _EmuHorusEngine)
    real, real_int = np.random.standard_normal, np.random.randint
    monkeypatch.setattr(np.random, "standard_normal", lambda n: real(min(n, 20000)))    # force the "stream too short" retries
    monkeypatch.setattr(np.random, "randint", lambda lo, hi=None, size=None, dtype=int: real_int(lo, hi, size=min(size, 50000), dtype=dtype))
    infra = infrastructure.Infrastructure(fl)
    jm = jobs.JobsManager(fl, jobs.JobQueueManager(fl, fl.trace_file))
    lm = log_manager.LogManager(str(tmp_path), fl)
    lm.init(infra)
    np.random.seed(params["seed"])
    stats = schedule.Scheduler(infra, jm, lm).start()
    assert stats.draws > 20000
    assert open(tmp_path / "job.csv", newline="").read() == job_csv
    assert open(tmp_path / "cluster.csv", newline="").read() == cluster_csv
