"""CPU-only differential test: the literal oracle (pinned to the reference) against the tight C
implementation of the engine's algorithm on many random clusters / traces, including the odd corners
(node counts that are not multiples of 32, 64-GPU nodes, cpu- or memory-bound nodes, over-sized jobs
that leak resources, gpu_per_container > 1, saturating arrival rates)."""
import numpy as np
import pytest


def _case(seed):
    from gpuschedule_b200 import capi, ingest, tracegen
    rng = np.random.default_rng(seed)
    G = int(rng.choice([1, 2, 4, 8, 8, 16, 64]))
    cluster = capi.make_cluster(num_switch=int(rng.integers(1, 4)), num_node_p_switch=int(rng.integers(1, 45)),
                                num_gpu_p_node=G, num_cpu_p_node=int(rng.choice([12, 24, 60, 128, 800])),
                                mem_p_node=int(rng.choice([60, 120, 300, 512, 4000])),
                                gpu_memory_capacity=int(rng.choice([16, 32])))
    gpc = int(rng.choice([1, 1, 2, 4])) if G >= 4 else 1
    choices = sorted(set(int(x) * gpc for x in rng.choice([1, 2, 3, 4, 8, 16, 40], size=4)))
    table = ingest.table_from_columns(tracegen.synth_columns(
        int(rng.integers(1, 400)), seed=5000 + seed, rate=float(rng.choice([0.2, 1.0, 4.0])), gpu_per_container=gpc,
        gpu_choices=choices, gpu_probs=rng.dirichlet(np.ones(len(choices))),
        max_mem_mib=int(rng.choice([16384, 17000, 33500]))))
    return cluster, table


@pytest.mark.parametrize("block", range(6))
def test_literal_oracle_equals_tight_cpu(block):
    import oracle
    for seed in range(block * 25, block * 25 + 25):
        cluster, table = _case(seed)
        ref = oracle.run_fifo(cluster, table)
        got = oracle.run_tight(cluster, table)
        assert got.ticks == ref.ticks and got.events == ref.events, seed
        assert got.rows.tobytes() == ref.rows.tobytes(), seed
        assert got.recs.tobytes() == ref.recs.tobytes(), seed
        assert np.array_equal(got.finish_order, ref.finish_order), seed
        assert np.array_equal(got.span_off, ref.span_off) and got.spans.tobytes() == ref.spans.tobytes(), seed
