"""Host-side policy parameters for the event-driven (Tiresias-style) policies.

`build_gittins_table` restates parse_job_dist / cal_r_gittins_index
(/root/reference/run_sim.py:1650-1708) with vectorised binary searches and prefix sums instead of the
O(n^2) generator scans; the arithmetic (Python round(), association) is kept.  The reference's
sample file yarn-gput1000.csv is not in the repository, so the sample is the trace's own
`run length x gpus`.  Pinned: tests/test_policy_golden.py compares the table with the
output of the reference's parse_job_dist executed verbatim (tests/golden/make_policy_golden.py).
"""
from __future__ import annotations

import sys

import numpy as np


def _py_round(x, nd):
    """Python's round(x, nd) for a float64 array: the value correctly rounded to nd decimals (ties on the EXACT binary
    value to even), then the nearest double.  rint(x * 10^nd) / 10^nd is that whenever x * 10^nd is not within rounding
    error of a half-way point -- the quotient of two exactly representable numbers is correctly rounded, like strtod of
    the decimal string --; the few entries that are go through Python's round()."""
    scale = 10.0 ** nd
    y = x * scale
    out = np.rint(y) / scale
    frac = np.abs(y - np.floor(y) - 0.5)
    risky = np.nonzero(~(frac > 8.0 * np.spacing(np.abs(y))))[0]               # x * 10^nd is within 1 ulp; also catches nan / inf
    for i in risky.tolist():
        out[i] = round(float(x[i]), nd)
    return out


def build_gittins_table(samples, delta=3250.0):
    """-> (data float64[n+1], index float64[n+1]); data sorted, last entries = (maxsize, 0.0)."""
    data = np.sort(np.asarray([int(x) for x in samples] if not isinstance(samples, np.ndarray) else samples, dtype=np.int64))
    num = len(data)
    if num == 0:
        return np.array([float(sys.maxsize)]), np.array([0.0])
    if int(data[-1]) * num >= 2 ** 53:                  # the sums below are exact in float64 only up to there
        raise ValueError("gittins sample too large")
    prefix = np.zeros(num + 1, dtype=np.int64)
    np.cumsum(data, out=prefix[1:])
    last = int(data[-1])
    # cal_r_gittins_index(job_data, a) for a = v - 1 of every sample v, all at once
    a = data - 1
    live = ~(a > last - 1)
    idx = np.searchsorted(data, a, side="right")       # first i with data[i] > a
    next_a = a + delta                                  # float, like the reference's a + delta
    idx_delta = np.where(next_a > last - 1, num - 1, np.searchsorted(data, next_a, side="right"))
    den = np.maximum(num - idx, 1).astype(np.float64)
    p = _py_round(((idx_delta - idx) * 1.0) / den, 5)
    e_sum = (prefix[idx_delta] - prefix[idx]).astype(np.float64) + (delta * (num - idx_delta))
    e = _py_round(e_sum / den, 5)
    with np.errstate(divide="ignore", invalid="ignore"):
        gi = _py_round(p * 1000000 / e, 4)
    if np.any(live & ~np.isfinite(gi)):
        raise ZeroDivisionError("float division by zero")       # what the reference raises for such a sample
    gi = np.where(live, gi, 0.0)
    return (np.concatenate([data.astype(np.float64), [float(sys.maxsize)]]),
            np.concatenate([gi, [0.0]]))


def gittins_samples(table):
    """run length (ticks) x gpus of every job of a JobTable."""
    need = np.maximum(1, np.ceil(table.duration)).astype(np.int64)
    return need * table.gpus.astype(np.int64)
