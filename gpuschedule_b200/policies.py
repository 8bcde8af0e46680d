"""Host-side policy parameters for the event-driven (Tiresias-style) policies.

`build_gittins_table` restates parse_job_dist / cal_r_gittins_index
(/root/reference/run_sim.py:1650-1708) with binary searches and prefix sums instead of the
O(n^2) generator scans; the arithmetic (Python round(), association) is kept.  The reference's
sample file yarn-gput1000.csv is not in the repository, so the sample is the trace's own
`run length x gpus`.  Pinned: tests/test_policy_golden.py compares the table with the
output of the reference's parse_job_dist executed verbatim (tests/golden/make_policy_golden.py).
"""
from __future__ import annotations

import bisect
import sys

import numpy as np


def build_gittins_table(samples, delta=3250.0):
    """-> (data float64[n+1], index float64[n+1]); data sorted, last entries = (maxsize, 0.0)."""
    data = sorted(int(x) for x in samples)
    num = len(data)
    if num == 0:
        return np.array([float(sys.maxsize)]), np.array([0.0])
    prefix = [0]
    for v in data:
        prefix.append(prefix[-1] + v)
    last = data[-1]

    def r_index(a):                                    # cal_r_gittins_index(job_data, a)
        if a > last - 1:
            return 0.0
        idx = bisect.bisect_right(data, a)             # first i with data[i] > a
        next_a = a + delta
        if next_a > last - 1:
            idx_delta = num - 1
        else:
            idx_delta = bisect.bisect_right(data, next_a)
        p = round(((idx_delta - idx) * 1.0) / (num - idx), 5)
        e_sum = (prefix[idx_delta] - prefix[idx]) + (delta * (num - idx_delta))
        e = round(e_sum / (num - idx), 5)
        return round(p * 1000000 / e, 4)

    gi = [r_index(int(v - 1)) for v in data]
    return (np.array([float(v) for v in data] + [float(sys.maxsize)], dtype=np.float64),
            np.array(gi + [0.0], dtype=np.float64))


def gittins_samples(table):
    """run length (ticks) x gpus of every job of a JobTable."""
    need = np.maximum(1, np.ceil(table.duration)).astype(np.int64)
    return need * table.gpus.astype(np.int64)
