"""Pins the CPU oracle (oracle/gsched_oracle.c) and the host formatting /
RNG-column replay against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Byte-for-byte: job.csv and all 13 columns of
cluster.csv."""
import pytest

from conftest import golden_cases, load_golden, render_outputs


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference_bytes(case):
    import oracle
    table, cluster, meta, job_csv, cluster_csv = load_golden(case)
    r = oracle.run_fifo(cluster, table)
    got_job, got_cluster = render_outputs(table, cluster, r.rows, r.recs, r.finish_order,
                                          r.span_off, r.spans, meta["numpy_seed"])
    assert got_job == job_csv
    got_lines = got_cluster.split("\r\n")
    exp_lines = cluster_csv.split("\r\n")
    assert len(got_lines) == len(exp_lines)
    for i, (a, b) in enumerate(zip(got_lines, exp_lines)):
        assert a == b, f"cluster.csv line {i}: {a!r} != {b!r}"


def test_kat0_known_answers():
    """The known-answer vector embedded in SURVEY.md 8(c)."""
    import oracle
    table, cluster, meta, _, _ = load_golden("kat0")
    r = oracle.run_fifo(cluster, table)
    assert r.ticks == 57
    got = [(table.label[j], int(r.recs[j]["start"]), int(r.recs[j]["end"])) for j in r.finish_order]
    assert got[:4] == [("0", 0, 11), ("1", 2, 15), ("2", 3, 22), ("3", 8, 23)]
    assert [table.label[j] for j in r.finish_order] == ["0", "1", "2", "3", "4", "5", "6", "8", "11", "7", "10", "9"]
    row16 = r.rows[15]
    assert (row16["now"], row16["idle_nodes"], row16["busy_nodes"], row16["busy_gpus"]) == (16, 122, 6, 41)
    # job 5 (16 GPUs) spans nodes 1,5,6 with 7+8+1 tasks
    j5 = table.label.index("5")
    sp = r.spans[r.span_off[j5]:r.span_off[j5 + 1]]
    assert [(int(s["node"]) + 1, int(s["ntasks"])) for s in sp] == [(1, 7), (5, 8), (6, 1)]


def test_net_cost_matches_reference_function():
    """oracle_net_cost == the reference's calculate_network_costs (network_service.py:3-39), bit for bit, on the
    400 vectors tests/golden/make_netcost_golden.py produced by calling the unmodified function."""
    import json
    import os
    import numpy as np
    import oracle
    from conftest import GOLDEN
    from gpuschedule_b200 import capi
    cases = json.load(open(os.path.join(GOLDEN, "netcost.json")))["cases"]
    assert len(cases) == 400
    for c in cases:
        cl = capi.make_cluster(4, 32, bandwidth=c["bandwidth"], internode_latency=c["latency"])
        got = oracle.net_cost(cl, np.array(c["node"], dtype=np.int32), np.array(c["is_ps"], dtype=np.uint8),
                              c["ps_count"], c["model_mb"], c["iterations"])
        assert got == float.fromhex(c["expected"]), c
