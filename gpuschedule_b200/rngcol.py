"""Host replay of the one stochastic cluster.csv column (avg_gpu_utilization).

The reference draws, every tick, one `np.random.normal(loc=util_avg,
scale=(util_max-util_avg)/2, size=1)` per busy device, walking nodes in id
order and devices 0..G-1, clips each at 100, accumulates sequentially and
divides by the GPU count (/root/reference/infra/device.py:48-54,
/root/reference/core/scheduling/schedule.py:103-120).  The draws come from
numpy's global legacy MT19937 stream, which is sequential by construction, so
this column stays on the host.  The engine supplies where every
job ran (gs_span records) and when (start/end); this module turns them into
(job, device, first row, last row) holdings, draws the standard-normal values
from numpy with ONE vectorised call per block of rows -- which consumes the
stream exactly like the per-device size=1 calls do -- and lets the library's
host helper (include/gsched.h gs_logcol_*) walk the ticks and busy devices in
the reference's order.  The text is produced with numpy's own dragon4 formatter.
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


def _intervals(n_rows, gpus_per_node, recs, span_off, spans):
    """(first row, last row, device key, job) of every (job, device) holding, sorted by first row.

    Row r (0-based) is the statistics row written with delta == r + 1; job j is
    counted there iff start_j <= r and (end_j < 0 or end_j > r + 1).
    """
    hold_job, hold_key = _holdings(recs, span_off, spans, gpus_per_node)
    start = recs["start"][hold_job].astype(np.int64)
    end = recs["end"][hold_job]
    first = start                                                   # first row counted
    last = np.where(end < 0, n_rows - 1, end.astype(np.int64) - 2)  # last row counted
    live = (start >= 0) & (last >= first) & (first < n_rows)
    hold_job, hold_key, first, last = hold_job[live], hold_key[live], first[live], np.minimum(last[live], n_rows - 1)
    order = np.argsort(first, kind="stable")
    return (np.ascontiguousarray(first[order]), np.ascontiguousarray(last[order]),
            np.ascontiguousarray(hold_key[order], dtype=np.int32), np.ascontiguousarray(hold_job[order], dtype=np.int32))


def _format_bracketed(values):
    """str(np.array([v])) for every v, fast: numpy prints a 1-element float64 array with the dragon4
    positional formatter (precision 8, unique, trim '.') unless the value calls for exponent form.  In the positional
    range that text is the value correctly rounded to 8 decimals with the trailing zeros removed -- what C's "%.8f"
    prints (both round the exact binary value, ties to even) -- because a shortest representation of 8 decimals or
    fewer is itself a multiple of 1e-8 within half an ulp of the value (tests/test_abi_and_host.py compares the two
    on millions of values, ties included)."""
    return [("[" + ("%.8f" % v).rstrip("0") + "]") if (1e-4 <= v < 1e8 or v == 0.0) else str(np.array([v]))   # exponent notation: leave it to numpy
            for v in (values.tolist() if isinstance(values, np.ndarray) else values)]


DRAWS_PER_CALL = 1 << 19        # standard-normal values drawn at a time: blocks that stay in the cache (fresh 100 MB arrays cost more in page faults than the draw itself)


def utilization_text(n_rows, n_nodes, gpus_per_node, table, recs, span_off, spans, rng=None):
    """Text of the avg_gpu_utilization column for rows 0..n_rows-1."""
    # numpy's legacy normal(loc, scale) is loc + scale * gauss() (legacy-distributions.c: legacy_normal); drawing the
    # standard values here and applying the affine map in gs_logcol_rows is the same two roundings (bit equality checked
    # on 21 M samples, tests/test_oracle_golden.py pins the bytes).  ONE vectorised draw per block of rows consumes the
    # stream exactly like the per-device size=1 calls do; the walk over ticks and busy devices is the library's
    # (csrc/gs_logcol.cpp), the text is numpy's own dragon4 formatter.
    from . import capi
    std_normal = np.random.standard_normal if rng is None else rng.standard_normal
    total = n_nodes * gpus_per_node
    loc = np.ascontiguousarray(table.util_avg, dtype=np.float64)
    scale = np.ascontiguousarray((table.util_max - table.util_avg) / 2, dtype=np.float64)
    first, last, key, job = _intervals(n_rows, gpus_per_node, recs, span_off, spans)
    out = []
    with capi.LogColumn(n_rows, total, first, last, key, job) as col:
        counts = col.counts()
        ends = np.cumsum(counts)
        r0 = 0
        while r0 < n_rows:
            base = int(ends[r0 - 1]) if r0 else 0
            r1 = max(r0 + 1, int(np.searchsorted(ends, base + DRAWS_PER_CALL, side="right")))
            r1 = min(r1, n_rows)
            z = std_normal(int(ends[r1 - 1]) - base)
            acc, n_arr = col.rows(r1, loc, scale, z)
            brack = _format_bracketed(acc / total)
            # no busy device: 0 / n -> float 0.0;  some un-clipped draw: numpy 1-element array;
            # every draw clipped at the int 100: plain Python numbers
            out.extend("0.0" if c == 0 else (bk if na > 0 else repr(float(int(ac) / total)))
                       for c, na, bk, ac in zip(counts[r0:r1].tolist(), n_arr.tolist(), brack, acc.tolist()))
            r0 = r1
    return out


def sampled_utilization_text(values, is_array):
    """Text of avg_gpu_utilization when the ENGINE sampled it (horus / gandiva: the samples are part of the
    scheduling decisions, so the column cannot be replayed afterwards).  A row whose sum saw an un-clipped
    sample is a 1-element numpy array in the reference and prints bracketed; otherwise it is a plain float."""
    bracketed = _format_bracketed(values)
    return [b if a else repr(float(v)) for b, a, v in zip(bracketed, np.asarray(is_array).tolist(), np.asarray(values).tolist())]
