"""oracle/tight_cpu.c (the engine's algorithm as tight single-thread C, used only as an extra CPU
yardstick in bench.py) must agree bit for bit with the pinned literal oracle and the fixtures."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden


@pytest.mark.parametrize("case", golden_cases())
def test_tight_cpu_matches_pinned_oracle_on_fixtures(case):
    import oracle
    table, cluster, _, _, _ = load_golden(case)
    if cluster.enable_network_costs:
        pytest.skip("the yardstick implements the plain fifo + yarn tick only (no network-cost branch)")
    ref = oracle.run_fifo(cluster, table)
    got = oracle.run_tight(cluster, table)
    assert got.ticks == ref.ticks and got.events == ref.events
    assert got.rows.tobytes() == ref.rows.tobytes()
    assert got.recs.tobytes() == ref.recs.tobytes()
    assert np.array_equal(got.finish_order, ref.finish_order)
    assert np.array_equal(got.span_off, ref.span_off) and got.spans.tobytes() == ref.spans.tobytes()


def test_tight_cpu_matches_oracle_seeded():
    import oracle
    from gpuschedule_b200 import capi, ingest, tracegen
    for seed, ckw in [(201, dict(num_switch=4, num_node_p_switch=32)), (202, dict(num_switch=1, num_node_p_switch=5)),
                      (203, dict(num_switch=2, num_node_p_switch=16, gpu_memory_capacity=16))]:
        cluster = capi.make_cluster(**ckw)
        table = ingest.table_from_columns(tracegen.synth_columns(3000, seed=seed, rate=0.9, max_mem_mib=17000))
        ref = oracle.run_fifo(cluster, table)
        got = oracle.run_tight(cluster, table)
        assert got.rows.tobytes() == ref.rows.tobytes() and got.recs.tobytes() == ref.recs.tobytes(), seed
