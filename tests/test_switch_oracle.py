"""SURVEY row a13, second half: the legacy switch-local yarn placement with parameter-server traffic accounting.
tests/golden/switch_yarn.json = the reference's own methods (infra/switch.py:38-167,190-206, infra/cluster.py:888-898)
executed verbatim under the stubs of tests/golden/make_switch_golden.py; oracle/switch_oracle.c must reproduce every
answer and every node table, doubles bit for bit -- including Python's round(x, 1) inside the traffic expression."""
import json
import os

import numpy as np

from conftest import GOLDEN


def load_switch_cases():
    with open(os.path.join(GOLDEN, "switch_yarn.json")) as f:
        return json.load(f)["cases"]


def check_case(case, place):
    """`place(num_gpu, total_size, ps_network)` -> (k, switch, node[], gpu[], cpu[], mem[], net[]) against the fixture"""
    for jd, ans in zip(case["jobs"], case["answers"]):
        k, sw, node, gpu, cpu, mem, net = place(jd["num_gpu"], jd["total_size"], jd["ps_network"])
        assert (k > 0) == ans["ok"], (case["name"], jd)
        if not ans["ok"]:
            continue
        assert sw == ans["switch"] and k == len(ans["nodes"]), (case["name"], jd)
        for i, nd in enumerate(ans["nodes"]):
            assert (int(node[i]), int(gpu[i]), int(cpu[i])) == (nd["id"], nd["num_gpu"], nd["num_cpu"]), (case["name"], jd, i)
            assert float(mem[i]).hex() == nd["mem"], (case["name"], jd, i)
            if nd["network"] is None:
                assert np.isnan(net[i])
            else:
                assert float(net[i]).hex() == nd["network"], (case["name"], jd, i, float(net[i]), float.fromhex(nd["network"]))


def test_round1_is_pythons_round():
    import oracle
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.uniform(-5000, 5000, 60000), rng.uniform(-3, 3, 20000), np.round(rng.uniform(-900, 900, 40000), 2),
                           np.arange(-200, 200) / 20.0, np.arange(-400, 400) / 40.0, [0.05, 0.15, 0.25, 0.35, 2.675, 1e11 + 0.25, -0.04, 0.0]])
    f = oracle.lib().switch_round1
    for v in vals.tolist():
        got, exp = f(v), round(v, 1)
        assert got == exp and np.signbit(got) == np.signbit(exp), (v, got, exp)


def test_switch_oracle_matches_the_reference_methods():
    import oracle
    for case in load_switch_cases():
        nodes = case["nodes"]
        cl = oracle.SwitchCluster(case["num_switch"], case["num_node_p_switch"], case["num_gpu_p_node"],
                                  [n["free_gpus"] for n in nodes], [n["free_cpus"] for n in nodes], [n["free_mem"] for n in nodes])
        check_case(case, cl.place)
        for i, a in enumerate(case["after"]):
            assert (int(cl.free_gpus[i]), int(cl.free_cpus[i])) == (a["free_gpus"], a["free_cpus"]), (case["name"], i)
            assert float(cl.free_mem[i]).hex() == a["free_mem"] and float(cl.net_in[i]).hex() == a["network_in"], (case["name"], i)
