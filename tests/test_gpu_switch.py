"""SURVEY row a13, second half on the GPU: gs_switch_yarn (legacy switch-local yarn placement + parameter-server traffic,
infra/switch.py:38-167, infra/cluster.py:888-898) against the fixtures made from the reference's own methods
(tests/golden/switch_yarn.json) -- every answer, every per-node traffic figure and every node table, doubles bit for
bit -- and against the pinned oracle on larger random clusters."""
import numpy as np
import pytest

from test_switch_oracle import check_case, load_switch_cases

pytestmark = pytest.mark.gpu


def _as_cluster(case):
    nodes = case["nodes"]
    return dict(num_switch=case["num_switch"], num_node_p_switch=case["num_node_p_switch"], num_gpu_p_node=case["num_gpu_p_node"],
                free_gpus=[n["free_gpus"] for n in nodes], free_cpus=[n["free_cpus"] for n in nodes], free_mem=[n["free_mem"] for n in nodes],
                jobs=[(j["num_gpu"], j["total_size"], j["ps_network"]) for j in case["jobs"]])


def test_switch_yarn_matches_the_reference_methods():
    from gpuschedule_b200 import capi
    cases = load_switch_cases()
    with capi.Engine(device=0, nsims=1) as eng:
        results = eng.switch_yarn([_as_cluster(c) for c in cases])          # all 60 clusters in one launch
    for case, (res, table) in zip(cases, results):
        it = iter(res)

        def place(g, total, ps):
            k, sw, sp = next(it)
            return k, sw, sp["node"], sp["num_gpu"], sp["num_cpu"], sp["mem"], sp["network"]
        check_case(case, place)
        for i, a in enumerate(case["after"]):
            assert (int(table["free_gpus"][i]), int(table["free_cpus"][i])) == (a["free_gpus"], a["free_cpus"]), (case["name"], i)
            assert float(table["free_mem"][i]).hex() == a["free_mem"] and float(table["net_in"][i]).hex() == a["network_in"], (case["name"], i)


def test_switch_yarn_matches_oracle_on_large_random_clusters():
    import oracle
    from gpuschedule_b200 import capi
    rng = np.random.default_rng(11)
    clusters = []
    for _ in range(40):
        S, P, G = int(rng.integers(1, 17)), int(rng.integers(20, 80)), int(rng.choice([4, 8, 16]))
        m = S * P
        used = rng.integers(0, G + 1, size=m) * (rng.random(m) < 0.5)
        jobs = []
        for _ in range(300):
            g = int(rng.choice([1, 2, 4, 8, 16, 24, 32, 40, 64, 100]))
            total = float(np.round(rng.uniform(1, 2000), 1))
            ps = [] if (g == 1 and rng.random() < 0.5) else [float(np.round(x, 1)) for x in rng.uniform(0, total / g * 2, size=g)]
            jobs.append((g, total, ps))
        clusters.append(dict(num_switch=S, num_node_p_switch=P, num_gpu_p_node=G, free_gpus=(G - used).tolist(),
                             free_cpus=(rng.integers(40, 200, size=m)).tolist(), free_mem=np.round(rng.uniform(50, 900, size=m), 1).tolist(), jobs=jobs))
    with capi.Engine(device=0, nsims=1) as eng:
        results = eng.switch_yarn(clusters)
    placed = 0
    for c, (res, table) in zip(clusters, results):
        ref = oracle.SwitchCluster(c["num_switch"], c["num_node_p_switch"], c["num_gpu_p_node"], c["free_gpus"], c["free_cpus"], c["free_mem"])
        for (g, total, ps), (k, sw, sp) in zip(c["jobs"], res):
            rk, rsw, rnode, rgpu, rcpu, rmem, rnet = ref.place(g, total, ps)
            assert k == rk and (k == 0 or sw == rsw)
            placed += k > 0
            assert np.array_equal(sp["node"], rnode) and np.array_equal(sp["num_gpu"], rgpu) and np.array_equal(sp["num_cpu"], rcpu)
            assert sp["mem"].tobytes() == rmem.tobytes() and sp["network"].tobytes() == rnet.tobytes()
        assert np.array_equal(table["free_gpus"], ref.free_gpus) and np.array_equal(table["free_cpus"], ref.free_cpus)
        assert table["free_mem"].tobytes() == ref.free_mem.tobytes() and table["net_in"].tobytes() == ref.net_in.tobytes()
    assert placed > 1000
