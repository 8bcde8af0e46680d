"""Host replay of the one stochastic cluster.csv column (avg_gpu_utilization).

The reference draws, every tick, one `np.random.normal(loc=util_avg,
scale=(util_max-util_avg)/2, size=1)` per busy device, walking nodes in id
order and devices 0..G-1, clips each at 100, accumulates sequentially and
divides by the GPU count (/root/reference/infra/device.py:48-54,
/root/reference/core/scheduling/schedule.py:103-120).  The draws come from
numpy's global legacy MT19937 stream, which is sequential by construction, so
this column stays on the host.  The engine supplies where every
job ran (gs_span records) and when (start/end); this module rebuilds the
per-tick ordered device list and replays the stream with ONE vectorised call
per chunk of rows, which consumes the stream exactly like the per-device
size=1 calls do.  Everything is vectorised numpy: a (rows x devices) owner grid
per chunk is a running sum of +owner / -owner marks at interval starts / stops, the
sequential per-row accumulation runs column-wise, and the text is produced with
numpy's own dragon4 formatter.
"""
from __future__ import annotations

import numpy as np


def _holdings(recs, span_off, spans, gpus_per_node):
    """One entry per (job, device): job index, device key = node * G + dev."""
    n = len(recs)
    if len(spans) == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z
    span_job = np.repeat(np.arange(n, dtype=np.int64), np.diff(span_off))
    masks = np.ascontiguousarray(spans["devmask"], dtype="<u8")
    bits = np.unpackbits(masks.view(np.uint8).reshape(-1, 8), axis=1, bitorder="little")   # (spans, 64)
    si, dev = np.nonzero(bits)
    return span_job[si], spans["node"][si].astype(np.int64) * gpus_per_node + dev.astype(np.int64)


def busy_job_stream(n_rows, n_nodes, gpus_per_node, recs, span_off, spans, chunk_rows=2048):
    """Yield (row_counts, jobs_in_draw_order) per chunk of rows.

    Row r (0-based) is the statistics row written with delta == r + 1; job j is
    counted there iff start_j <= r and (end_j < 0 or end_j > r + 1).
    """
    width = n_nodes * gpus_per_node
    hold_job, hold_key = _holdings(recs, span_off, spans, gpus_per_node)
    start = recs["start"][hold_job].astype(np.int64)
    end = recs["end"][hold_job]
    keep = start >= 0
    hold_job, hold_key, start, end = hold_job[keep], hold_key[keep], start[keep], end[keep]
    first = start                                                   # first row counted
    last = np.where(end < 0, n_rows - 1, end.astype(np.int64) - 2)  # last row counted
    live = last >= first
    hold_job, hold_key, first, last = hold_job[live], hold_key[live], first[live], last[live]
    order = np.argsort(first, kind="stable")
    hold_job, hold_key, first, last = hold_job[order], hold_key[order], first[order], last[order]
    lo_ptr = 0
    active = np.zeros(0, dtype=np.int64)                            # holdings still open
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        hi_ptr = int(np.searchsorted(first, r1, side="left"))
        cand = np.concatenate([active, np.arange(lo_ptr, hi_ptr, dtype=np.int64)])
        lo_ptr = hi_ptr
        cand = cand[last[cand] >= r0]
        rows_n = r1 - r0
        # +(job+1) where a holding starts (or continues into the chunk), -(job+1) on the first row after it;
        # devices host one job at a time, so a running sum along each device's row is the owner (+1), 0 = idle.
        # The grid is kept device-major so that the running sum walks contiguous memory, then transposed once.
        ukeys, kidx = np.unique(hold_key[cand], return_inverse=True)    # only devices that host something in this chunk
        marks = np.zeros((len(ukeys), rows_n + 1), dtype=np.int32)
        a = np.maximum(first[cand], r0) - r0
        b = np.minimum(last[cand], r1 - 1) - r0 + 1                 # first row after the holding (<= rows_n)
        val = (hold_job[cand] + 1).astype(np.int32)
        np.add.at(marks, (kidx, a), val)
        np.add.at(marks, (kidx, b), -val)
        np.cumsum(marks, axis=1, out=marks)
        filled = np.ascontiguousarray(marks[:, :rows_n].T)            # (rows, devices in id order)
        busy = filled > 0
        active = cand[last[cand] >= r1]
        yield busy.sum(axis=1), filled[busy].astype(np.int64) - 1


def _format_bracketed(values):
    """str(np.array([v])) for every v, fast: numpy prints a 1-element float64 array with the dragon4
    positional formatter (precision 8, unique, trim '.') unless the value calls for exponent form."""
    out = []
    fmt = np.format_float_positional
    for v in values:
        if v != 0.0 and (v < 1e-4 or v >= 1e8):
            out.append(str(np.array([v])))                          # exponent notation: leave it to numpy
        else:
            out.append("[" + fmt(v, precision=8, unique=True, fractional=True, trim=".") + "]")
    return out


def utilization_text(n_rows, n_nodes, gpus_per_node, table, recs, span_off, spans, rng=None):
    """Text of the avg_gpu_utilization column for rows 0..n_rows-1."""
    # numpy's legacy normal(loc, scale) is loc + scale * gauss() (legacy-distributions.c: legacy_normal); drawing the
    # standard values and applying the affine map in numpy is the same two roundings and 3x faster than broadcasting
    # loc / scale through the generator (bit equality checked on 21 M samples, tests/test_oracle_golden.py pins the bytes)
    std_normal = np.random.standard_normal if rng is None else rng.standard_normal
    total = n_nodes * gpus_per_node
    out = []
    loc_all = table.util_avg
    scale_all = (table.util_max - table.util_avg) / 2
    for counts, jobs in busy_job_stream(n_rows, n_nodes, gpus_per_node, recs, span_off, spans):
        if len(jobs):
            draw = loc_all[jobs] + scale_all[jobs] * std_normal(len(jobs))
        else:
            draw = np.zeros(0)
        clipped = draw >= 100.0          # min(100, x) returns the int 100 unless x < 100
        vals = np.where(clipped, 100.0, draw)
        off = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        # sequential per-row accumulation (Python float additions, left to right): the draws of a chunk go into a
        # (rows x widest row) matrix padded with 0.0 -- x + 0.0 == x -- and are added column by column
        acc = np.zeros(len(counts), dtype=np.float64)
        if len(vals):
            row_of = np.repeat(np.arange(len(counts)), counts)
            pos_of = np.arange(len(vals)) - off[row_of]
            mat = np.zeros((int(counts.max()), len(counts)), dtype=np.float64)
            mat[pos_of, row_of] = vals
            for k in range(mat.shape[0]):
                acc += mat[k]
        csum = np.concatenate([[0], np.cumsum(~clipped)])
        n_arr = csum[off[1:]] - csum[off[:-1]]                      # un-clipped draws per row (numpy arrays)
        frac = acc / total
        brack = _format_bracketed(frac)
        # no busy device: 0 / n -> float 0.0;  some un-clipped draw: numpy 1-element array;
        # every draw clipped at the int 100: plain Python numbers
        out.extend("0.0" if c == 0 else (bk if na > 0 else repr(float(int(ac) / total)))
                   for c, na, bk, ac in zip(counts.tolist(), n_arr.tolist(), brack, acc.tolist()))
    return out


def sampled_utilization_text(values, is_array):
    """Text of avg_gpu_utilization when the ENGINE sampled it (horus / gandiva: the samples are part of the
    scheduling decisions, so the column cannot be replayed afterwards).  A row whose sum saw an un-clipped
    sample is a 1-element numpy array in the reference and prints bracketed; otherwise it is a plain float."""
    bracketed = _format_bracketed(values)
    return [b if a else repr(float(v)) for b, a, v in zip(bracketed, np.asarray(is_array).tolist(), np.asarray(values).tolist())]
