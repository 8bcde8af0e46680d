"""Host replay of the one stochastic cluster.csv column (avg_gpu_utilization).

The reference draws, every tick, one `np.random.normal(loc=util_avg,
scale=(util_max-util_avg)/2, size=1)` per busy device, walking nodes in id
order and devices 0..G-1, clips each at 100, accumulates sequentially and
divides by the GPU count (/root/reference/infra/device.py:48-54,
/root/reference/core/scheduling/schedule.py:103-120).  The draws come from
numpy's global legacy MT19937 stream, which is sequential by construction, so
this column stays on the host.  The engine supplies where every
job ran (gs_span records) and when (start/end); this module rebuilds the
per-tick ordered device list and replays the stream with ONE vectorised call,
which consumes the stream exactly like the per-device size=1 calls do.
"""
from __future__ import annotations

import numpy as np


def busy_job_stream(n_rows, n_nodes, gpus_per_node, recs, span_off, spans, chunk_rows=2048):
    """Yield (row_counts, jobs_in_draw_order) per chunk of rows.

    Row r (0-based) is the statistics row written with delta == r + 1; job j is
    counted there iff start_j <= r and (end_j < 0 or end_j > r + 1).
    """
    G = gpus_per_node
    width = n_nodes * G
    start = recs["start"]
    end = recs["end"]
    started = np.nonzero(start >= 0)[0]
    # one entry per (job, device): key = node * G + dev
    hold_job, hold_key = [], []
    for j in started:
        for s in spans[span_off[j]:span_off[j + 1]]:
            mask = int(s["devmask"])
            base = int(s["node"]) * G
            d = 0
            while mask:
                if mask & 1:
                    hold_job.append(j)
                    hold_key.append(base + d)
                mask >>= 1
                d += 1
    hold_job = np.asarray(hold_job, dtype=np.int64)
    hold_key = np.asarray(hold_key, dtype=np.int64)
    first = start[hold_job].astype(np.int64)                       # first row counted
    last = np.where(end[hold_job] < 0, n_rows - 1, end[hold_job].astype(np.int64) - 2)
    order = np.argsort(first, kind="stable")
    hold_job, hold_key, first, last = hold_job[order], hold_key[order], first[order], last[order]
    lo_ptr = 0
    active = np.zeros(0, dtype=np.int64)                           # indices into hold_* still open
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        hi_ptr = int(np.searchsorted(first, r1, side="left"))
        cand = np.concatenate([active, np.arange(lo_ptr, hi_ptr, dtype=np.int64)])
        lo_ptr = hi_ptr
        cand = cand[last[cand] >= r0]
        grid = np.full((r1 - r0, width), -1, dtype=np.int64)
        for h in cand:
            a = max(int(first[h]), r0) - r0
            b = min(int(last[h]), r1 - 1) - r0 + 1
            if b > a:
                grid[a:b, hold_key[h]] = hold_job[h]
        active = cand[last[cand] >= r1]
        busy = grid >= 0
        yield busy.sum(axis=1), grid[busy]


def utilization_text(n_rows, n_nodes, gpus_per_node, table, recs, span_off, spans, rng=None):
    """Text of the avg_gpu_utilization column for rows 0..n_rows-1."""
    normal = np.random.normal if rng is None else rng.normal
    total = n_nodes * gpus_per_node
    out = []
    for counts, jobs in busy_job_stream(n_rows, n_nodes, gpus_per_node, recs, span_off, spans):
        if len(jobs):
            loc = table.util_avg[jobs]
            scale = (table.util_max[jobs] - table.util_avg[jobs]) / 2
            draw = normal(loc=loc, scale=scale)
        else:
            draw = np.zeros(0)
        clipped = draw >= 100.0          # min(100, x) returns the int 100 unless x < 100
        vals = np.where(clipped, 100.0, draw)
        off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        acc = np.zeros(len(counts), dtype=np.float64)
        for k in range(int(counts.max()) if len(counts) else 0):   # sequential per-row accumulation
            sel = np.nonzero(counts > k)[0]
            acc[sel] = acc[sel] + vals[off[sel] + k]
        n_arr = np.add.reduceat(np.concatenate([~clipped, [False]]).astype(np.int64),
                                np.minimum(off[:-1], len(clipped)))
        n_arr = np.where(counts > 0, n_arr, 0)
        for i in range(len(counts)):
            if counts[i] == 0:
                out.append("0.0")                                   # 0 / n  -> float 0.0
            elif n_arr[i] > 0:
                out.append(str(np.array([acc[i] / total])))         # numpy 1-element array
            else:
                out.append(repr(float(int(acc[i]) / total)))        # every draw clipped: plain ints
    return out
