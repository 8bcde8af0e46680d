"""oracle/tight2_cpu.c = the event-stepped algorithm of the CUDA fifo engine (gs_tick2.cuh) as scalar C, emitting the
same compact records (gs_evrow / gs_qrow / gs_job_run / start-ordered spans) in resumable windows.  Decoded with the
package's own decoders it must agree bit for bit with the pinned oracle -- this checks, on a box without a GPU, the
algorithm the kernel implements (tick jumping, register-resident head, front-pushed wheel, blocked-head shortcut,
window resume) AND log_manager.expand_rows / expand_jobs / group_spans."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden


def _same(ref, got, tag):
    assert got.ticks == ref.ticks, tag
    assert got.events == ref.events, tag
    if got.rows.tobytes() != ref.rows.tobytes():
        for i in range(min(len(got.rows), len(ref.rows))):
            assert got.rows[i].tobytes() == ref.rows[i].tobytes(), f"{tag} row {i}: {got.rows[i]} != {ref.rows[i]}"
    assert len(got.rows) == len(ref.rows), tag
    assert got.recs.tobytes() == ref.recs.tobytes(), tag
    assert np.array_equal(got.finish_order, ref.finish_order), tag
    assert np.array_equal(got.span_off, ref.span_off) and got.spans.tobytes() == ref.spans.tobytes(), tag
    assert got.evals == ref.evals, tag


@pytest.mark.parametrize("case", golden_cases())
def test_tight2_matches_pinned_oracle_on_fixtures(case):
    import oracle
    table, cluster, _, _, _ = load_golden(case)
    if cluster.enable_network_costs:
        pytest.skip("the yardstick implements the plain fifo + yarn tick only (no network-cost branch)")
    ref = oracle.run_fifo(cluster, table)
    t2 = oracle.Tight2(cluster, table)
    _same(ref, t2.run_all(), case)
    _same(ref, t2.run_all(max_ticks=7), case + " 7-tick windows")
    _same(ref, t2.run_all(cap_a=1, cap_b=1), case + " one record per window")


def test_tight2_matches_oracle_on_random_cases():
    import oracle
    from test_cpu_differential import _case
    for seed in range(500, 560):
        cluster, table = _case(seed)
        ref = oracle.run_fifo(cluster, table)
        t2 = oracle.Tight2(cluster, table)
        _same(ref, t2.run_all(), f"seed {seed}")
        _same(ref, t2.run_all(max_ticks=1 + seed % 13, cap_a=1 + seed % 5, cap_b=1 + seed % 3), f"seed {seed} windows")
