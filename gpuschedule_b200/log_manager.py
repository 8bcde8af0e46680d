"""Output surface: cluster.csv (one row per tick), job.csv (finish order) and the
four header-only CSVs, byte-identical to the reference's writer
(/root/reference/log_manager.py:5-155): same headers, `\\r\\n` line ends (Python
csv default, quirk Q16), same value types so that str() prints the same text
(`[0.05556386]` arrays, `nan`, `0` vs `19.0`; quirks Q14, Q15, Q22).

The per-object API (`LogInfo`, `step_cluster`, `jcts`) is kept for drop-in use;
the engine uses the batch forms (`write_cluster_rows`, `write_job_rows`) that
format whole runs at once from the integer aggregates the device produced.
"""
from __future__ import annotations

import os
import time

import numpy as np

CLUSTER_HEADER = ["delta", "num_idle_nodes", "num_busy_nodes", "num_busy_gpus",
                  "num_idle_gpus", "avg_gpu_utilization", "avg_gpu_memory_allocated",
                  "avg_pending_time", "median_pending_time", "max_pending_time",
                  "num_running_jobs", "num_queuing_jobs", "num_finish_jobs"]
JOB_HEADER = ["job_id", "num_gpu", "submit_time", "start_time", "end_time",
              "original_duration", "actual_duration", "jct", "preempt"]
EOL = "\r\n"

# numpy dtype mirroring include/gsched.h: gs_tick_row / gs_job_rec / gs_span
ROW_DTYPE = np.dtype([("now", "<i4"), ("idle_nodes", "<i4"), ("busy_nodes", "<i4"),
                      ("busy_gpus", "<i4"), ("idle_gpus", "<i4"), ("running", "<i4"),
                      ("queued", "<i4"), ("finished", "<i4"), ("mem_busy_bytes", "<i8"),
                      ("pend_sum", "<i8"), ("pend_max", "<i4"), ("pend_med_lo", "<i4"),
                      ("pend_med_hi", "<i4"), ("reserved", "<i4")])
JOB_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("jct", "<i4"),
                      ("preempt", "<i4"), ("duration", "<f8")])
SPAN_DTYPE = np.dtype([("node", "<i4"), ("ntasks", "<i4"), ("devmask", "<u8")])
assert ROW_DTYPE.itemsize == 64 and JOB_DTYPE.itemsize == 24 and SPAN_DTYPE.itemsize == 16
# compact records of the fifo engine: gs_evrow / gs_qrow / gs_job_start / gs_cspan
EVROW_DTYPE = np.dtype([("now", "<i4"), ("queued", "<i4"), ("finished", "<i4"), ("busy_gpus", "<u2"), ("running", "<u2"),
                        ("mem_busy_bytes", "<i8")])
QROW_DTYPE = np.dtype([("now", "<i4"), ("oldest_arrive", "<i4"), ("med_lo_arrive", "<i4"), ("med_hi_arrive", "<i4"),
                       ("arrive_sum", "<i8")])
NODEEV_DTYPE = np.dtype([("now", "<i4"), ("busy_nodes", "<i4")])
JOBRUN_DTYPE = np.dtype([("start", "<i4")])                     # gs_job_start: the start tick, -1 = never started
CSPAN_DTYPE = np.dtype([("where", "<u4"), ("devmask", "<u4")])  # gs_cspan (clusters with at most 32 GPUs per node)
assert (EVROW_DTYPE.itemsize, QROW_DTYPE.itemsize, NODEEV_DTYPE.itemsize, JOBRUN_DTYPE.itemsize, CSPAN_DTYPE.itemsize) == (24, 24, 8, 4, 8)
SPAN_FIRST = 0x80000000


def expand_rows(ev, qr, nodeev, row_first, ticks, n_nodes, gpus_per_node):
    """gs_tick_row of every tick of a window from its compact records (include/gsched.h: gs_evrow / gs_qrow / gs_nodeev):
    record k describes the rows `now_k` .. `now_(k+1) - 1`; on those only `delta` and the pending
    statistics move, linearly with the tick (pending = now - arrival, jobs_manager.py:72-87)."""
    count = int(ticks - row_first)
    rows = np.zeros(count, dtype=ROW_DTYPE)
    if count == 0:
        return rows
    now = np.arange(row_first + 1, ticks + 1, dtype=np.int64)
    starts = ev["now"].astype(np.int64)
    gaps = np.diff(np.append(starts, ticks + 1))
    k = np.repeat(np.arange(len(ev)), gaps)
    e = ev[k]
    q = e["queued"].astype(np.int64)
    busy = e["busy_gpus"].astype(np.int32)
    nodes = nodeev["busy_nodes"][np.searchsorted(nodeev["now"], ev["now"], side="right") - 1][k]   # last node event at or before the record
    rows["now"] = now
    rows["idle_nodes"] = n_nodes - nodes
    rows["busy_nodes"] = nodes
    rows["busy_gpus"] = busy
    rows["idle_gpus"] = n_nodes * gpus_per_node - busy
    rows["running"] = e["running"].astype(np.int32)
    rows["queued"] = e["queued"]
    rows["finished"] = e["finished"]
    rows["mem_busy_bytes"] = e["mem_busy_bytes"]
    has = q > 0
    if has.any():
        b = qr[np.searchsorted(qr["now"], e["now"][has])]           # the queue record taken on the same tick
        v = now[has]
        rows["pend_sum"][has] = q[has] * v - b["arrive_sum"]
        rows["pend_max"][has] = v - b["oldest_arrive"]
        rows["pend_med_lo"][has] = v - b["med_lo_arrive"]
        rows["pend_med_hi"][has] = v - b["med_hi_arrive"]
    return rows


def expand_jobs(job_run, admitted, duration_in, duration_out=None):
    """gs_job_rec view of the compact per-job result (the start tick).  fifo never preempts: the run length is
    max(1, ceil(Job.get_duration())) ticks (quirk Q11; job.py:206-210), end = start + run, jct = run,
    preempt = migration_count = 1 (Q12); never started: -1, -1, 0, 0."""
    n = len(job_run)
    recs = np.zeros(n, dtype=JOB_DTYPE)
    started = (np.arange(n) < admitted) & (job_run["start"] >= 0)
    dur = np.asarray(duration_in, dtype=np.float64)
    if duration_out is not None:
        dur = np.where(started, duration_out, dur)
    eff = np.maximum(dur, duration_in)
    run = np.where(np.ceil(eff) < 1.0, 1, np.minimum(np.ceil(eff), 2147483647.0)).astype(np.int64).astype(np.int32)
    recs["start"] = np.where(started, job_run["start"], -1)
    recs["end"] = np.where(started, job_run["start"] + run, -1)
    recs["jct"] = np.where(started, run, 0)
    recs["preempt"] = started.astype(np.int32)
    recs["duration"] = dur
    return recs


def widen_spans(pool):
    """gs_cspan records -> gs_span records (the first-of-job flag moves into bit 31 of ntasks)"""
    if pool.dtype == SPAN_DTYPE:
        return pool
    w = pool["where"]
    out = np.zeros(len(pool), dtype=SPAN_DTYPE)
    out["node"] = (w & 0xfffff).astype(np.int32)
    out["ntasks"] = ((((w >> 20) & 0x3f) + 1) | (w & np.uint32(SPAN_FIRST))).astype(np.uint32).view(np.int32)
    out["devmask"] = pool["devmask"].astype(np.uint64)
    return out


def group_spans(job_run, admitted, pool):
    """(span_off[n+1], spans grouped by job) from the start-ordered span pool of gs_fetch_compact: the k-th
    record flagged SPAN_FIRST opens the job with the k-th smallest start tick (start ticks are unique)."""
    pool = widen_spans(pool)
    n = len(job_run)
    started = np.nonzero((np.arange(n) < admitted) & (job_run["start"] >= 0))[0]
    order = started[np.argsort(job_run["start"][started], kind="stable")]
    first = (pool["ntasks"].view(np.uint32) & SPAN_FIRST) != 0
    owner = order[np.cumsum(first) - 1] if len(pool) else np.zeros(0, dtype=np.int64)
    cnt = np.bincount(owner, minlength=n) if len(pool) else np.zeros(n, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    out = pool[np.argsort(owner, kind="stable")].copy()
    out["ntasks"] = (out["ntasks"].view(np.uint32) & ~np.uint32(SPAN_FIRST)).astype(np.int32)
    return off, out


class LogInfo:
    """Same fields as the reference record (log_manager.py:5-30)."""

    def __init__(self, num_idle_nodes, num_busy_nodes, num_busy_gpus, num_idle_gpus,
                 avg_gpu_utilization, avg_gpu_memory_allocated, avg_pending_time,
                 median_pending_time, max_pending_time, num_running_jobs,
                 num_queuing_jobs, num_finish_jobs):
        self.idle_ns = num_idle_nodes
        self.busy_ns = num_busy_nodes
        self.busy_gs = num_busy_gpus
        self.idle_gs = num_idle_gpus
        self.avg_g_utils = avg_gpu_utilization
        self.avg_g_mem = avg_gpu_memory_allocated
        self.avg_pending = avg_pending_time
        self.median_pending = median_pending_time
        self.max_pending = max_pending_time
        self.num_running_jobs = num_running_jobs
        self.num_queuing_jobs = num_queuing_jobs
        self.num_finish_jobs = num_finish_jobs

    def fields(self, delta):
        return [delta, self.idle_ns, self.busy_ns, self.busy_gs, self.idle_gs,
                self.avg_g_utils, self.avg_g_mem, self.avg_pending, self.median_pending,
                self.max_pending, self.num_running_jobs, self.num_queuing_jobs,
                self.num_finish_jobs]


def _line(values):
    return ",".join(v if isinstance(v, str) else str(v) for v in values) + EOL


def pending_columns(rows):
    """avg / median / max pending text from the integer aggregates, using the
    reference's float expressions (jobs_manager.py:72-87).  Vectorised: only rows with a
    non-empty queue need float formatting (str(np.float64) and repr(float) print the same text)."""
    q = rows["queued"]
    n = len(rows)
    avg_t, med_t, max_t = ["0.0"] * n, ["nan"] * n, ["0"] * n
    nz = np.nonzero(q > 0)[0]
    if len(nz):
        avg = rows["pend_sum"][nz].astype(np.float64) / (q[nz].astype(np.float64) + 1e-9)
        med = (rows["pend_med_lo"][nz].astype(np.float64) + rows["pend_med_hi"][nz].astype(np.float64)) / 2.0
        mx = rows["pend_max"][nz].astype(np.float64)
        for i, a, m, x in zip(nz.tolist(), avg.tolist(), med.tolist(), mx.tolist()):
            avg_t[i] = repr(a); med_t[i] = repr(m); max_t[i] = repr(x)
    return avg_t, med_t, max_t


def memory_column(rows, total_cap_mib):
    """sum(min(cap, MiB)) / sum(cap) (schedule.py:115,121; device.py:56-62)."""
    vals = (rows["mem_busy_bytes"].astype(np.float64) / 1048576.0) / total_cap_mib
    busy = rows["busy_gpus"] > 0
    return [repr(v) if b else "0.0" for v, b in zip(vals.tolist(), busy.tolist())]


def int_text(col):
    """str() of every entry of an integer column.  The statistics columns are counters with a small range, so the text
    of each distinct value is made once and gathered (numpy's int -> str conversion costs ~10x a gather)."""
    col = np.asarray(col)
    if len(col) == 0:
        return []
    lo, hi = int(col.min()), int(col.max())
    if hi - lo > 4 * len(col) + 4096:
        return col.astype(str).tolist()
    table = np.array([str(v) for v in range(lo, hi + 1)], dtype=object)
    return table[col.astype(np.int64) - lo].tolist()


def cluster_static_columns(rows, total_cap_mib):
    """Text of every cluster.csv column but the sampled one: (five leading columns, memory, avg / median / max pending,
    three trailing columns).  Independent of the utilisation column, so a caller can prepare it meanwhile."""
    avg_t, med_t, max_t = pending_columns(rows)
    mem_t = memory_column(rows, total_cap_mib)
    cols = [int_text(rows[k]) for k in ("now", "idle_nodes", "busy_nodes", "busy_gpus", "idle_gpus")]
    tail = [int_text(rows[k]) for k in ("running", "queued", "finished")]
    return cols, mem_t, (avg_t, med_t, max_t), tail


def cluster_lines(rows, util_text, total_cap_mib, static=None):
    cols, mem_t, (avg_t, med_t, max_t), tail = static if static is not None else cluster_static_columns(rows, total_cap_mib)
    return [",".join(t) for t in zip(cols[0], cols[1], cols[2], cols[3], cols[4], util_text, mem_t, avg_t, med_t, max_t,
                                     tail[0], tail[1], tail[2])]


def job_lines(table, recs, finish_order):
    """One line per finished job in finish order (log_manager.py:137-153)."""
    order = np.asarray(finish_order, dtype=np.int64)
    dur = recs["duration"][order]
    actual = np.maximum(table.duration[order], dur)              # Job.get_duration job.py:206-210
    label, gtext = table.label, table.num_gpu_text
    return [",".join((label[j], gtext[j], sb, st, en, repr(d), repr(a), jc, pr))
            for j, sb, st, en, d, a, jc, pr in zip(order.tolist(), int_text(table.submit[order]), int_text(recs["start"][order]),
                                                   int_text(recs["end"][order]), dur.tolist(), actual.tolist(),
                                                   int_text(recs["jct"][order]), int_text(recs["preempt"][order]))]


def horus_job_lines(table, recs, finish_order):
    """job.csv lines from gs_horus_job_rec records: Job.duration and Job.get_duration() differ once a task has
    carried an interference penalty (jobs_manager.py:189-201), so both come from the engine."""
    order = np.asarray(finish_order, dtype=np.int64)
    label, gtext = table.label, table.num_gpu_text
    return [",".join((label[j], gtext[j], str(sb), str(st), str(en), repr(o), repr(a), str(jc), str(pr)))
            for j, sb, st, en, o, a, jc, pr in zip(order.tolist(), table.submit[order].tolist(), recs["start"][order].tolist(),
                                                   recs["end"][order].tolist(), recs["original"][order].tolist(),
                                                   recs["actual"][order].tolist(), recs["jct"][order].tolist(),
                                                   recs["preempt"][order].tolist())]


class LogManager:
    def __init__(self, log_path, flags):
        self.log_path = log_path
        self.flags = flags
        self.is_count = getattr(flags, "scheme", "yarn") == "count"
        self.cluster_stats_header = list(CLUSTER_HEADER)
        self.job_stats_header = list(JOB_HEADER)

    def init(self, infrastructure):
        p = self.log_path
        self.log_cluster = os.path.join(p, "cluster.csv")
        self.log_job = os.path.join(p, "job.csv")
        with open(self.log_cluster, "w", newline="") as f:
            f.write(_line(self.cluster_stats_header))
        if not self.is_count:
            n_nodes = len(infrastructure.nodes)
            self.log_cpu = os.path.join(p, "cpu.csv")
            self.log_gpu = os.path.join(p, "gpu.csv")
            self.log_network = os.path.join(p, "network.csv")
            self.log_mem = os.path.join(p, "memory.csv")
            with open(self.log_cpu, "w", newline="") as f:
                f.write(_line(["time"] + ["cpu%d" % i for i in range(n_nodes)]))
            with open(self.log_gpu, "w", newline="") as f:
                f.write(_line(["time"] + ["gpu%d" % i for i in range(infrastructure.get_total_gpus())]))
            with open(self.log_mem, "w", newline="") as f:
                f.write(_line(["time", "max", "99th", "95th", "med"]))
            with open(self.log_network, "w", newline="") as f:
                titles = ["time"]
                for i in range(n_nodes):
                    titles += ["in%d" % i, "out%d" % i]
                f.write(_line(titles))
        with open(self.log_job, "w", newline="") as f:
            f.write(_line(self.job_stats_header))
        assert os.path.exists(self.log_cluster)

    # ---- per-object API (drop-in)
    def step_cluster(self, loginfo, delta):
        with open(self.log_cluster, "a", newline="") as f:
            f.write(_line(loginfo.fields(delta)))

    def jcts(self, finished_jobs):
        assert len(finished_jobs) > 0, ValueError("No finished jobs")
        with open(self.log_job, "a", newline="") as f:
            for _, j in finished_jobs.items():
                f.write(_line([j.job_id, j.gpus, j.submit_time, j.start_time, j.end_time,
                               j.duration, j.get_duration(), j.time_processed(), j.migration_count]))
        time.sleep(1)

    # ---- batch API (engine)
    def write_cluster_rows(self, rows, util_text, total_cap_mib, static=None):
        lines = cluster_lines(rows, util_text, total_cap_mib, static)
        with open(self.log_cluster, "a", newline="") as f:
            if lines:
                f.write(EOL.join(lines) + EOL)

    def write_horus_job_rows(self, table, recs, finish_order):
        assert len(finish_order) > 0, ValueError("No finished jobs")
        with open(self.log_job, "a", newline="") as f:
            f.write(EOL.join(horus_job_lines(table, recs, finish_order)) + EOL)

    def write_job_rows(self, table, recs, finish_order, lines=None):
        assert len(finish_order) > 0, ValueError("No finished jobs")
        lines = job_lines(table, recs, finish_order) if lines is None else lines
        with open(self.log_job, "a", newline="") as f:
            f.write(EOL.join(lines) + EOL)


def render_horus_job_csv(table, recs, finish_order):
    body = horus_job_lines(table, recs, finish_order)
    return _line(JOB_HEADER) + (EOL.join(body) + EOL if body else "")


def render_cluster_csv(rows, util_text, total_cap_mib):
    """Whole cluster.csv as text (header + rows)."""
    body = cluster_lines(rows, util_text, total_cap_mib)
    return _line(CLUSTER_HEADER) + (EOL.join(body) + EOL if body else "")


def render_job_csv(table, recs, finish_order):
    body = job_lines(table, recs, finish_order)
    return _line(JOB_HEADER) + (EOL.join(body) + EOL if body else "")
