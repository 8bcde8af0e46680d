"""Trace ingest: CSV -> admission-ordered structure-of-arrays job table.

Host-side (pandas) on purpose: the reference orders rows with pandas' default
*unstable* quicksort (/root/reference/core/jobs/job_generator.py:185), so tie
order can only be reproduced by issuing the same pandas calls in the same
sequence.  Everything after that is turned into flat integer
/ float64 arrays for the device engine:

  arrive_tick = first integer tick with normalized_time <= tick     (job_generator.py:203)
  submit      = int(normalized_time)                                (job.py:93)
  duration    = minutes * scale_factor                              (jobs_manager.py:234, schedule.py:187)
  tasks       = used_gpus // gpu_per_container                      (job.py:96-98)
  mem_bytes   = memory_max (integral bytes; MiB = bytes / 2^20)     (jobs_manager.py:236, util.py:21-29)
"""
from __future__ import annotations

import logging
import os
import sys
from dataclasses import dataclass, field

import numpy as np

REQUIRED = ["type", "normalized_time", "minutes", "gpu_per_container",
            "gpu_utilization_avg", "gpu_utilization_max", "memory_max",
            "memory_avg", "used_gpus"]


@dataclass
class JobTable:
    n: int
    label: list                      # job_id == CSV row label (jobs_manager.py:233-234)
    num_gpu_text: list               # how LogManager prints Job.gpus (dtype of the column, Q22)
    arrive_tick: np.ndarray          # int32
    submit: np.ndarray               # int32
    gpus: np.ndarray                 # int32
    gpu_per_task: np.ndarray         # int32
    duration: np.ndarray             # float64
    mem_bytes: np.ndarray            # int64
    util_avg: np.ndarray             # float64 (host only: RNG column)
    util_max: np.ndarray             # float64
    model_mb: np.ndarray | None = None
    iterations: np.ndarray | None = None
    ps_count: np.ndarray | None = None
    extra: dict = field(default_factory=dict)

    @property
    def tasks(self):
        return self.gpus // self.gpu_per_task

    def packed(self):
        """The trace as 32-byte gs_jobin records (cached): what gs_load_trace_packed consumes."""
        if "packed" not in self.extra:
            from .capi import JOBIN_DTYPE
            a = np.empty(self.n, dtype=JOBIN_DTYPE)
            a["arrive_tick"] = self.arrive_tick
            a["gpus"] = self.gpus
            a["gpu_per_task"] = self.gpu_per_task
            a["ps_count"] = 0 if self.ps_count is None else self.ps_count
            a["mem_bytes"] = self.mem_bytes
            a["duration"] = self.duration
            self.extra["packed"] = a
        return self.extra["packed"]

    def task_offsets(self):
        off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(self.tasks, out=off[1:])
        return off


class JobTraceReader:
    """Same constructor/`prepare_jobs` contract as the reference reader
    (/root/reference/core/jobs/job_generator.py:165-193), including its error
    behaviour: a missing or unreadable file logs an error and exits 1."""

    def __init__(self, file_path):
        import pandas as pd
        if not os.path.exists(file_path):
            logging.error(f"file: {file_path} not exist")
            sys.exit(1)
        try:
            self.trace_df = pd.read_csv(file_path)
        except Exception:
            logging.error(f"unable to read the file, assumed it was csv. but got {file_path}")
            sys.exit(1)
        self.prepared = False

    def prepare_jobs(self):
        # identical call sequence to job_generator.py:183-189 (filter, sort, dropna, shift, scale)
        df = self.trace_df
        df = df[df["type"] == "noninteractive"]
        df.sort_values(by="normalized_time", inplace=True)
        df.dropna(inplace=True)
        df["normalized_time"] -= df["normalized_time"].min()
        df["normalized_time"] /= 10000
        self.trace_df = df
        self.prepared = True
        return self

    def remaining_jobs(self):
        return len(self.trace_df)

    def table(self, scale_factor=0.5) -> JobTable:
        if not self.prepared:
            self.prepare_jobs()
        return table_from_frame(self.trace_df, scale_factor)


def _as_i32(name, values):
    v = np.asarray(values, dtype=np.float64)
    if v.size and (np.any(v != np.floor(v)) or np.any(np.abs(v) >= 2 ** 31)):
        raise ValueError(f"column {name!r} must hold integers below 2^31")
    return v.astype(np.int32)


def table_from_frame(df, scale_factor=0.5) -> JobTable:
    """`df` is the prepared frame (normalized_time already in ticks, rows in
    admission order)."""
    n = len(df)
    nt = df["normalized_time"].to_numpy(dtype=np.float64)
    if n and (np.any(nt < 0) or np.any(nt >= 2 ** 31 - 2)):
        raise ValueError("normalized_time out of range")
    arrive = np.ceil(nt).astype(np.int32)
    submit = nt.astype(np.int64).astype(np.int32)           # int() truncates
    gpus_col = df["used_gpus"]
    gpus = _as_i32("used_gpus", gpus_col.to_numpy())
    gpc = _as_i32("gpu_per_container", df["gpu_per_container"].to_numpy())
    if n and (np.any(gpc <= 0) or np.any(gpus < gpc)):
        # zero-task jobs crash the reference (StopIteration at node.py:117, quirk Q24)
        raise ValueError("used_gpus must be a positive multiple of gpu_per_container")
    if n and np.any(gpus % gpc != 0):
        raise ValueError("used_gpus must be a multiple of gpu_per_container")
    if gpus_col.dtype.kind == "f":
        num_gpu_text = [repr(float(x)) for x in gpus_col.to_numpy()]
    else:
        num_gpu_text = [str(int(x)) for x in gpus_col.to_numpy()]
    duration = df["minutes"].to_numpy(dtype=np.float64) * scale_factor
    if n and np.any(duration >= 2 ** 30):
        raise ValueError("duration too large")
    mem = df["memory_max"].to_numpy(dtype=np.float64)
    if n and (np.any(mem != np.floor(mem)) or np.any(mem < 0) or np.any(mem >= 2 ** 53)):
        raise ValueError("memory_max must be integral bytes (bit-exact memory column, quirk Q23)")
    t = JobTable(
        n=n, label=[str(x) for x in df.index], num_gpu_text=num_gpu_text,
        arrive_tick=arrive, submit=submit, gpus=gpus, gpu_per_task=gpc,
        duration=np.ascontiguousarray(duration), mem_bytes=mem.astype(np.int64),
        util_avg=df["gpu_utilization_avg"].to_numpy(dtype=np.float64),
        util_max=df["gpu_utilization_max"].to_numpy(dtype=np.float64))
    if "memory_avg" in df.columns:      # Job.gpu_mem_avg in MiB (jobs_manager.py:237): a k-means feature (core/jobs/utils.py:4-22)
        t.extra["mem_avg_mib"] = df["memory_avg"].to_numpy(dtype=np.float64) / 1024 / 1024
    if "model_name" in df.columns or "model_size" in df.columns:
        from .model_factory import model_size_mb
        if "model_size" in df.columns:
            t.model_mb = df["model_size"].to_numpy(dtype=np.float64)
        else:
            t.model_mb = np.array([model_size_mb(x) for x in df["model_name"]], dtype=np.float64)
    if "iterations" in df.columns:
        t.iterations = df["iterations"].to_numpy(dtype=np.float64)
    if "ps_count" in df.columns:
        t.ps_count = _as_i32("ps_count", df["ps_count"].to_numpy())
    return t


CACHE_FORMAT = 1
_ARRAY_FIELDS = ("arrive_tick", "submit", "gpus", "gpu_per_task", "duration", "mem_bytes", "util_avg", "util_max")
_OPTIONAL_FIELDS = ("model_mb", "iterations", "ps_count")


def _cache_path(file_path, scale_factor, cache_dir):
    """One file per (trace file contents as far as stat() tells, scale factor, pandas / numpy versions: the row order is
    whatever THOSE versions' sort produces, see the module docstring)."""
    import hashlib
    from importlib import metadata
    st = os.stat(file_path)
    key = "|".join([os.path.abspath(file_path), str(st.st_size), str(st.st_mtime_ns), repr(float(scale_factor)),
                    metadata.version("pandas"), metadata.version("numpy"), str(CACHE_FORMAT)])
    return os.path.join(cache_dir, hashlib.sha1(key.encode()).hexdigest() + ".npz")


def load_table(file_path, scale_factor=0.5, cache_dir=None) -> JobTable:
    """The prepared JobTable of a trace file.  Parsing goes through pandas (JobTraceReader) -- whose import alone costs a
    second -- so a sweep that replays one trace many times (execute.py) keeps the parsed table under `cache_dir` and
    later runs read that instead; any change of the file (size, mtime), of the scale factor or of the pandas / numpy
    versions misses.  cache_dir=None: no cache."""
    path = None
    if cache_dir is not None and os.path.exists(file_path):
        try:
            path = _cache_path(file_path, scale_factor, cache_dir)
            if os.path.exists(path):
                with np.load(path, allow_pickle=False) as z:
                    t = JobTable(n=int(z["n"]), label=z["label"].tolist(), num_gpu_text=z["num_gpu_text"].tolist(),
                                 **{k: z[k] for k in _ARRAY_FIELDS})
                    for k in _OPTIONAL_FIELDS:
                        if k in z.files:
                            setattr(t, k, z[k])
                    if "mem_avg_mib" in z.files:
                        t.extra["mem_avg_mib"] = z["mem_avg_mib"]
                return t
        except Exception as exc:                            # noqa: BLE001 - a bad cache file is a miss, never an error
            logging.warning("trace cache %s unusable (%s): parsing the trace", path, exc)
    t = JobTraceReader(file_path).prepare_jobs().table(scale_factor)
    if path is not None:
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            arrays = {k: getattr(t, k) for k in _ARRAY_FIELDS}
            arrays.update({k: getattr(t, k) for k in _OPTIONAL_FIELDS if getattr(t, k) is not None})
            if "mem_avg_mib" in t.extra:
                arrays["mem_avg_mib"] = t.extra["mem_avg_mib"]
            tmp = path + ".%d.tmp.npz" % os.getpid()
            np.savez(tmp, n=np.int64(t.n), label=np.array(t.label, dtype=str), num_gpu_text=np.array(t.num_gpu_text, dtype=str), **arrays)
            os.replace(tmp, path)
        except OSError as exc:
            logging.warning("trace cache not written (%s)", exc)
    return t


def table_from_columns(cols: dict, scale_factor=0.5) -> JobTable:
    """Synthetic columns (tracegen.synth_columns) -> table through the same
    pandas path as a CSV on disk, without touching the filesystem."""
    import pandas as pd
    r = JobTraceReader.__new__(JobTraceReader)
    r.trace_df = pd.DataFrame(cols)
    r.prepared = False
    return r.table(scale_factor)
