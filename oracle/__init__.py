"""TEST INFRASTRUCTURE -- the CPU checker for libgsched.so.

ctypes wrapper around oracle/liboracle.so (oracle/gsched_oracle.c), the plain-C
restatement of the reference's live fifo+yarn tick loop.  It may be imported
only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` arm.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gpuschedule_b200.capi import JOBREQ_DTYPE, NODE_DTYPE, GsCluster
from gpuschedule_b200.log_manager import JOB_DTYPE, ROW_DTYPE, SPAN_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "gsched_oracle.c"), os.path.join(_HERE, "policy_oracle.c"),
            os.path.join(_HERE, "tight_cpu.c"), os.path.join(_HERE, "tight2_cpu.c"), os.path.join(_HERE, "horus_oracle.c"),
            os.path.join(_HERE, "switch_oracle.c")]
    hdr = os.path.join(os.path.dirname(_HERE), "include", "gsched.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(x) for x in srcs + [hdr])):
        return LIB_PATH
    subprocess.run(["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-shared",
                    "-o", LIB_PATH] + srcs + ["-lm"], check=True, cwd=_HERE)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_run_fifo.restype = C.c_int64
        _lib.oracle_run_policy.restype = C.c_int64
        _lib.tight_run_fifo.restype = C.c_int64
        _lib.tight2_create.restype = C.c_void_p
        _lib.tight2_jobs.restype = C.c_void_p
        _lib.tight2_finish_order.restype = C.c_void_p
        _lib.tight2_spans.restype = C.c_void_p
        _lib.tight2_run.restype = C.c_int
        _lib.switch_round1.restype = C.c_double
        _lib.switch_round1.argtypes = [C.c_double]
        _lib.switch_yarn_place.restype = C.c_int
        _lib.oracle_run_horus.restype = C.c_int64
        _lib.oracle_place_one.restype = C.c_int
        _lib.oracle_net_cost.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleResult:
    pass


def run_fifo(cluster: GsCluster, table, rows_cap=None, want_spans=True):
    """Run the restated Scheduler.start() on a JobTable; returns an OracleResult
    with the same arrays the engine's fetch_* calls give."""
    n = table.n
    if rows_cap is None:
        rows_cap = int(table.arrive_tick[-1] if n else 0) + int(np.ceil(table.duration).sum() if n else 0) + n + 16
        rows_cap = min(rows_cap, 1 << 26)
    rows = np.zeros(rows_cap, dtype=ROW_DTYPE)
    recs = np.zeros(max(n, 1), dtype=JOB_DTYPE)
    order = np.zeros(max(n, 1), dtype=np.int32)
    task_off = table.task_offsets()
    task_node = np.full(max(int(task_off[-1]), 1), -1, dtype=np.int32)
    task_mask = np.zeros(max(int(task_off[-1]), 1), dtype=np.uint64)
    nfin = C.c_int64(0)
    events = C.c_int64(0)
    evals = C.c_int64(0)
    arr = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
    a, g, c = arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32)
    d, m = arr(table.duration, np.float64), arr(table.mem_bytes, np.int64)
    mm, it, ps = arr(table.model_mb, np.float64), arr(table.iterations, np.float64), arr(table.ps_count, np.int32)
    ticks = lib().oracle_run_fifo(C.byref(cluster), C.c_int64(n), _p(a), _p(g), _p(c), _p(d), _p(m),
                                  _p(mm), _p(it), _p(ps), _p(rows), C.c_int64(rows_cap), _p(recs),
                                  _p(order), C.byref(nfin), _p(task_off), _p(task_node), _p(task_mask),
                                  C.byref(events), C.byref(evals))
    if ticks < 0:
        raise RuntimeError(f"oracle_run_fifo failed: {ticks}")
    r = OracleResult()
    r.ticks = int(ticks)
    r.rows = rows[:ticks]
    r.recs = recs[:n]
    r.finish_order = order[:nfin.value]
    r.events = events.value
    r.evals = evals.value
    r.task_off, r.task_node, r.task_mask = task_off, task_node, task_mask
    if want_spans:
        r.span_off, r.spans = spans_from_tasks(n, task_off, task_node, task_mask)
    return r


def spans_from_tasks(n, task_off, task_node, task_mask):
    """Collapse the per-task placement log into gs_span records (job, node)."""
    off = np.zeros(n + 1, dtype=np.int64)
    out = []
    for j in range(n):
        a, b = int(task_off[j]), int(task_off[j + 1])
        cur = None
        for t in range(a, b):
            nd = int(task_node[t])
            if nd < 0:
                break
            if cur is not None and cur[0] == nd:
                cur[1] += 1
                cur[2] |= int(task_mask[t])
            else:
                cur = [nd, 1, int(task_mask[t])]
                out.append(cur)
        off[j + 1] = len(out)
    spans = np.zeros(len(out), dtype=SPAN_DTYPE)
    for i, (nd, k, mk) in enumerate(out):
        spans[i] = (nd, k, mk)
    return off, spans


def place_one(cluster: GsCluster, nodes, job):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
    jr = np.zeros(1, dtype=JOBREQ_DTYPE)
    jr[0] = job
    tasks = int(jr[0]["gpus"]) // int(jr[0]["gpu_per_task"])
    task_node = np.full(max(tasks, 1), -1, dtype=np.int32)
    first = C.c_int32(-1)
    used = C.c_int32(0)
    ok = lib().oracle_place_one(C.byref(cluster), _p(nodes), C.c_int32(len(nodes)), _p(jr),
                                C.byref(first), C.byref(used), _p(task_node))
    return bool(ok), first.value, used.value, task_node[:tasks]


def net_cost(cluster: GsCluster, task_node, is_ps, ps_count, model_mb, iterations):
    task_node = np.ascontiguousarray(task_node, dtype=np.int32)
    is_ps = None if is_ps is None else np.ascontiguousarray(is_ps, dtype=np.uint8)
    return float(lib().oracle_net_cost(C.byref(cluster), C.c_int32(len(task_node)), _p(task_node),
                                       _p(is_ps), C.c_int32(int(ps_count)), C.c_double(float(model_mb)),
                                       C.c_double(float(iterations))))


def run_policy(cluster: GsCluster, policy, table, rows_cap=None):
    """Event-driven policies (sjf / dlas / dlas-gpu / gittins), oracle/policy_oracle.c."""
    n = table.n
    if rows_cap is None:
        rows_cap = 8 * n + 64
    rows = np.zeros(rows_cap, dtype=ROW_DTYPE)
    recs = np.zeros(max(n, 1), dtype=JOB_DTYPE)
    order = np.zeros(max(n, 1), dtype=np.int32)
    nfin = C.c_int64(0)
    events = C.c_int64(0)
    arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    a, g, c = arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32)
    d, m = arr(table.duration, np.float64), arr(table.mem_bytes, np.int64)
    ticks = lib().oracle_run_policy(C.byref(cluster), C.byref(policy), C.c_int64(n), _p(a), _p(g), _p(c), _p(d), _p(m),
                                    _p(rows), C.c_int64(rows_cap), _p(recs), _p(order), C.byref(nfin), C.byref(events))
    if ticks < 0:
        raise RuntimeError(f"oracle_run_policy failed: {ticks}")
    r = OracleResult()
    r.ticks = int(ticks)
    r.rows = rows[:ticks]
    r.recs = recs[:n]
    r.finish_order = order[:nfin.value]
    r.events = events.value
    return r


class TightRunner:
    """Pre-allocated buffers + one bare C call: what bench.py times for the `cpu_tight` yardstick."""

    def __init__(self, cluster: GsCluster, table):
        n = table.n
        self.cluster, self.n = cluster, n
        self.rows_cap = int(table.arrive_tick[-1]) + 2 * int(np.ceil(table.duration.max())) + 4096
        self.rows = np.zeros(self.rows_cap, dtype=ROW_DTYPE)
        self.recs = np.zeros(n, dtype=JOB_DTYPE)
        self.order = np.zeros(n, dtype=np.int32)
        m = cluster.num_switch * cluster.num_node_p_switch
        self.cap = int(np.minimum(table.tasks, m).sum()) + 1
        self.spans = np.zeros(self.cap, dtype=SPAN_DTYPE)
        self.sfirst = np.zeros(n, dtype=np.int64)
        self.scnt = np.zeros(n, dtype=np.int32)
        arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        self.cols = (arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32),
                     arr(table.duration, np.float64), arr(table.mem_bytes, np.int64))
        self.fn = lib().tight_run_fifo

    def run(self):
        nfin, events = C.c_int64(0), C.c_int64(0)
        a, g, c, d, mm = self.cols
        ticks = self.fn(C.byref(self.cluster), C.c_int64(self.n), _p(a), _p(g), _p(c), _p(d), _p(mm), _p(self.rows),
                        C.c_int64(self.rows_cap), _p(self.recs), _p(self.order), C.byref(nfin), _p(self.spans),
                        C.c_int64(self.cap), _p(self.sfirst), _p(self.scnt), C.byref(events))
        if ticks < 0:
            raise RuntimeError(f"tight_run_fifo failed: {ticks}")
        return int(ticks), int(events.value)


def run_tight(cluster: GsCluster, table, rows_cap=None):
    """oracle/tight_cpu.c: the engine's own algorithm as tight single-thread C (no network cost)."""
    n = table.n
    if rows_cap is None:
        rows_cap = int(table.arrive_tick[-1] if n else 0) + 2 * int(np.ceil(table.duration.max()) if n else 0) + n + 4096
    rows = np.zeros(rows_cap, dtype=ROW_DTYPE)
    recs = np.zeros(max(n, 1), dtype=JOB_DTYPE)
    order = np.zeros(max(n, 1), dtype=np.int32)
    m = cluster.num_switch * cluster.num_node_p_switch
    cap = int(np.minimum(table.tasks, m).sum()) + 1
    spans = np.zeros(cap, dtype=SPAN_DTYPE)
    sfirst = np.zeros(max(n, 1), dtype=np.int64)
    scnt = np.zeros(max(n, 1), dtype=np.int32)
    nfin, events = C.c_int64(0), C.c_int64(0)
    arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    a, g, c = arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32)
    d, mm = arr(table.duration, np.float64), arr(table.mem_bytes, np.int64)
    ticks = lib().tight_run_fifo(C.byref(cluster), C.c_int64(n), _p(a), _p(g), _p(c), _p(d), _p(mm), _p(rows),
                                 C.c_int64(rows_cap), _p(recs), _p(order), C.byref(nfin), _p(spans), C.c_int64(cap),
                                 _p(sfirst), _p(scnt), C.byref(events))
    if ticks < 0:
        raise RuntimeError(f"tight_run_fifo failed: {ticks}")
    r = OracleResult()
    r.ticks, r.rows, r.recs = int(ticks), rows[:ticks], recs[:n]
    r.finish_order, r.events = order[:nfin.value], events.value
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(scnt[:n], out=off[1:])
    idx = np.concatenate([np.arange(f, f + k) for f, k in zip(sfirst[:n], scnt[:n])]) if n and off[-1] else np.zeros(0, dtype=np.int64)
    r.span_off, r.spans = off, spans[idx.astype(np.int64)] if len(idx) else spans[:0]
    return r


HORUS_REC_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("jct", "<i4"), ("preempt", "<i4"),
                            ("original", "<f8"), ("actual", "<f8")])
# Which score function runs is decided by the SCHEDULE name: Scheduler._schedule passes self.schedule as the `scheme`
# argument of the scheduling algorithm, which hands it to horus_placement, which indexes score_fn with it
# (schedule.py:47, algorithm.py:9-13,58,196).  --scheme only selects the placement routine, and horus / horus+ /
# gandiva all map to horus_placement (algorithm.py:182-187).  schedule "fifo" with such a scheme raises KeyError there.
HORUS_SCHEMES = ("horus", "horus+", "gandiva", "yarn")       # yarn: ms_yarn_placement under the same schedulers
HORUS_SCORE_OF_SCHEDULE = {"horus": 0, "horus+": 0, "gandiva": 1}
HORUS_SCHEDULES = {"fifo": 0, "horus": 1, "horus+": 2, "gandiva": 3}


def run_horus(cluster: GsCluster, table, scheme="horus", schedule="horus", num_buffer=5, num_queue=1, seed=0, rows_cap=None):
    """The utilisation-aware live paths (horus / horus+ / gandiva), oracle/horus_oracle.c.  `seed` is the value
    numpy.random.seed() was given before the reference run being checked."""
    n = table.n
    if rows_cap is None:
        rows_cap = min(int(table.arrive_tick[-1] if n else 0) + 2 * int(np.ceil(table.duration).sum() if n else 0) + 8 * n + 64, 1 << 24)
    rows = np.zeros(rows_cap, dtype=ROW_DTYPE)
    util = np.zeros(rows_cap, dtype=np.float64)
    util_arr = np.zeros(rows_cap, dtype=np.uint8)
    recs = np.zeros(max(n, 1), dtype=HORUS_REC_DTYPE)
    order = np.zeros(max(n, 1), dtype=np.int32)
    nfin, events, draws = C.c_int64(0), C.c_int64(0), C.c_uint64(0)
    arr = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt)
    a, g, c = arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32)
    d, m = arr(table.duration, np.float64), arr(table.mem_bytes, np.int64)
    ma = arr(table.extra.get("mem_avg_mib"), np.float64)
    ua, um = arr(table.util_avg, np.float64), arr(table.util_max, np.float64)
    if scheme not in HORUS_SCHEMES or schedule not in HORUS_SCORE_OF_SCHEDULE:
        raise KeyError(f"scheme {scheme!r} / schedule {schedule!r}: the reference raises here too (score_fn[schedule])")
    ticks = lib().oracle_run_horus(C.byref(cluster), C.c_int32(HORUS_SCORE_OF_SCHEDULE[schedule] | (256 if scheme == "yarn" else 0)), C.c_int32(HORUS_SCHEDULES[schedule]),
                                   C.c_int32(num_buffer), C.c_int32(num_queue), C.c_uint32(seed), C.c_int64(n),
                                   _p(a), _p(g), _p(c), _p(d), _p(m), _p(ma), _p(ua), _p(um),
                                   _p(rows), _p(util), _p(util_arr), C.c_int64(rows_cap), _p(recs), _p(order),
                                   C.byref(nfin), C.byref(events), C.byref(draws))
    if ticks < 0:
        raise RuntimeError(f"oracle_run_horus failed: {ticks}")
    r = OracleResult()
    r.ticks = int(ticks)
    r.rows, r.util, r.util_is_array = rows[:ticks], util[:ticks], util_arr[:ticks]
    r.recs = recs[:n]
    r.finish_order = order[:nfin.value]
    r.events, r.draws = events.value, draws.value
    return r


class Tight2:
    """oracle/tight2_cpu.c: the event-stepped algorithm of the CUDA fifo engine as single-thread C, producing the same
    compact records in resumable windows.  run_all() decodes them with the PACKAGE's decoders (log_manager.expand_rows /
    expand_jobs / group_spans), so a comparison with the pinned oracle checks algorithm and decoders together."""

    def __init__(self, cluster: GsCluster, table, span_cap=0):
        from gpuschedule_b200.capi import GsWindowInfo
        from gpuschedule_b200.log_manager import EVROW_DTYPE, JOBRUN_DTYPE, NODEEV_DTYPE, QROW_DTYPE
        self._w, self._dt = GsWindowInfo, (EVROW_DTYPE, QROW_DTYPE, JOBRUN_DTYPE)
        self.cluster, self.table, self.n = cluster, table, table.n
        arr = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        self.cols = (arr(table.arrive_tick, np.int32), arr(table.gpus, np.int32), arr(table.gpu_per_task, np.int32),
                     arr(table.duration, np.float64), arr(table.mem_bytes, np.int64))
        a, g, c, d, mm = self.cols
        self.h = C.c_void_p(lib().tight2_create(C.byref(cluster), C.c_int64(self.n), _p(a), _p(g), _p(c), _p(d), _p(mm), C.c_int64(span_cap)))
        # at most one job starts per tick, so a saturated run lasts at least n ticks
        worst = max(int(table.arrive_tick[-1] if self.n else 0), self.n) + 2 * int(np.ceil(table.duration.max()) if self.n else 0) + 4096
        self.cap = worst
        self.ev = np.zeros(worst, dtype=EVROW_DTYPE)
        self.qr = np.zeros(worst, dtype=QROW_DTYPE)
        self.ne = np.zeros(cluster.num_switch * cluster.num_node_p_switch + 2, dtype=NODEEV_DTYPE)

    def __del__(self):
        if getattr(self, "h", None):
            lib().tight2_free(self.h)
            self.h = None

    def restart(self):
        lib().tight2_restart(self.h)

    def run_window(self, max_ticks=0, cap_a=None, cap_b=None):
        """one window -> (status, GsWindowInfo, events, evals, done); records are in self.ev / self.qr"""
        nev, nq, nne = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        rc = lib().tight2_run(self.h, C.c_int64(max_ticks), C.c_int64(self.cap if cap_a is None else cap_a),
                              C.c_int64(self.cap if cap_b is None else cap_b), _p(self.ev), _p(self.qr), _p(self.ne),
                              C.byref(nev), C.byref(nq), C.byref(nne))
        w = self._w()
        events, evals, done = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        lib().tight2_info(self.h, C.byref(w), nev, nq, nne, C.byref(events), C.byref(evals), C.byref(done))
        return rc, w, events.value, evals.value, done.value

    def run(self):
        """bare timing loop body: one full run into the preallocated record buffers -> (ticks, events)"""
        self.restart()
        rc, w, events, _, done = self.run_window()
        if rc != 0 or not done:
            raise RuntimeError(f"tight2_run failed: {rc}")
        return int(w.ticks), int(events)

    def run_all(self, max_ticks=0, cap_a=None, cap_b=None):
        """run to the end in windows; returns an OracleResult in the legacy row / record formats"""
        from gpuschedule_b200 import log_manager as lm
        self.restart()
        m, g = self.cluster.num_switch * self.cluster.num_node_p_switch, self.cluster.num_gpu_p_node
        parts = []
        while True:
            rc, w, events, evals, done = self.run_window(max_ticks, cap_a, cap_b)
            if rc != 0:
                raise RuntimeError(f"tight2_run failed: {rc}")
            parts.append(lm.expand_rows(self.ev[:w.ev_rows], self.qr[:w.q_rows], self.ne[:w.node_events], w.row_first, w.ticks, m, g))
            if done or self.n == 0:
                break
        r = OracleResult()
        n = self.n
        both = np.ctypeslib.as_array(C.cast(lib().tight2_jobs(self.h), C.POINTER(C.c_int32)), shape=(max(n, 1) * 2,)).reshape(-1, 2)[:n]
        jobs = np.zeros(n, dtype=self._dt[2])                 # the engine's compact per-job result: the start tick only
        jobs["start"] = both[:, 0]
        r.ticks = int(w.ticks)
        r.rows = np.concatenate(parts) if parts else np.zeros(0, dtype=ROW_DTYPE)
        r.recs = lm.expand_jobs(jobs, int(w.admitted), self.cols[3])
        nf = int(w.finished)
        r.finish_order = np.ctypeslib.as_array(C.cast(lib().tight2_finish_order(self.h), C.POINTER(C.c_int32)), shape=(max(nf, 1),))[:nf].copy()
        ns = int(w.spans_used)
        pool = np.ctypeslib.as_array(C.cast(lib().tight2_spans(self.h), C.POINTER(C.c_uint8)), shape=(max(ns, 1) * 16,)).view(SPAN_DTYPE)[:ns].copy()
        if g <= 32:                                           # the engine's 8-byte records for such clusters
            from gpuschedule_b200.log_manager import CSPAN_DTYPE, SPAN_FIRST
            c8 = np.zeros(ns, dtype=CSPAN_DTYPE)
            nt = pool["ntasks"].view(np.uint32)
            c8["where"] = pool["node"].astype(np.uint32) | (((nt & 0x7fffffff) - 1) << 20) | (nt & np.uint32(SPAN_FIRST))
            c8["devmask"] = pool["devmask"].astype(np.uint32)
            pool = c8
        r.span_off, r.spans = lm.group_spans(jobs, int(w.admitted), pool)
        r.events, r.evals, r.job_run, r.pool = events, evals, jobs, pool
        return r


SWITCH_MEM = (5.0, 8.0, 0.2)          # worker_mem, ps_mem, p_w_mem (core/models.py:24-26)


class SwitchCluster:
    """Node tables of the legacy switch-local yarn placement (oracle/switch_oracle.c); place() is one
    _Cluster.ms_yarn_placement call and mutates the tables on success."""

    def __init__(self, S, P, G, free_gpus, free_cpus, free_mem):
        self.S, self.P, self.G = S, P, G
        self.free_gpus = np.ascontiguousarray(free_gpus, dtype=np.int32).copy()
        self.free_cpus = np.ascontiguousarray(free_cpus, dtype=np.int32).copy()
        self.free_mem = np.ascontiguousarray(free_mem, dtype=np.float64).copy()
        self.net_in = np.zeros(S * P, dtype=np.float64)

    def place(self, num_gpu, model_size, ps_network):
        ps = np.ascontiguousarray(ps_network, dtype=np.float64)
        cap = self.P + 1
        sw = C.c_int32(-1)
        node, gpu, cpu = (np.zeros(cap, dtype=np.int32) for _ in range(3))
        mem, net = np.zeros(cap), np.zeros(cap)
        k = lib().switch_yarn_place(C.c_int(self.S), C.c_int(self.P), C.c_int(self.G), _p(self.free_gpus), _p(self.free_cpus),
                                    _p(self.free_mem), _p(self.net_in), C.c_int(int(num_gpu)), C.c_double(float(model_size)),
                                    _p(ps), C.c_int(len(ps)), C.c_double(SWITCH_MEM[0]), C.c_double(SWITCH_MEM[1]),
                                    C.c_double(SWITCH_MEM[2]), C.byref(sw), _p(node), _p(gpu), _p(cpu), _p(mem), _p(net))
        return k, sw.value, node[:k], gpu[:k], cpu[:k], mem[:k], net[:k]
