#!/usr/bin/env python
"""Golden fixtures for the utilisation-aware live paths (horus / horus+ / gandiva), made by running the
UNMODIFIED reference (build container only):   python tests/golden/make_horus_golden.py [case ...]

Same harness as make_golden.py (runpy of /root/reference/run_sim.py from a scratch CWD under
numpy.random.seed(SEED)); only --scheme / --schedule / --num_queue / --num_buffer differ.  These paths draw
from numpy's global stream inside the scheduling decisions, so the seed is part of the fixture.
Writes tests/golden/horus_<case>/{trace.csv,flags.json,job.csv,cluster.csv}.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402
from gpuschedule_b200 import tracegen  # noqa: E402


def frame(n, seed, rate, **kw):
    return tracegen.synth_frame(n, seed=seed, rate=rate, **kw).drop(columns=["model"])


CASES = {
    # name: (frame builder, flags)
    "horus_small": (lambda: frame(60, 61, 0.8), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8,
                    _scheme="horus", _schedule="horus", num_buffer=5)),
    "horus_racks": (lambda: frame(120, 62, 1.5, gpu_choices=[1, 2, 4, 8, 12, 16], gpu_probs=[.3, .2, .2, .15, .1, .05]),
                    dict(num_switch=3, num_node_p_switch=3, num_gpu_p_node=4, _scheme="horus", _schedule="horus", num_buffer=3)),
    "horus_buf1": (lambda: frame(100, 63, 2.0), dict(num_switch=1, num_node_p_switch=3, num_gpu_p_node=8,
                   _scheme="horus", _schedule="horus", num_buffer=1)),
    "gandiva_small": (lambda: frame(60, 64, 0.8), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8,
                      _scheme="gandiva", _schedule="gandiva", num_buffer=1)),
    "gandiva_slice": (lambda: frame(90, 65, 1.5, gpu_choices=[1, 2, 4, 8, 12], gpu_probs=[.3, .3, .2, .1, .1]),
                      dict(num_switch=2, num_node_p_switch=2, num_gpu_p_node=8, _scheme="gandiva", _schedule="gandiva", num_buffer=1)),
    "horusplus_k3": (lambda: frame(80, 66, 1.0), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8,
                     _scheme="horus+", _schedule="horus+", num_queue=3, num_buffer=15)),
    # score function and scheduler follow the --schedule name, the placement routine the --scheme name (see
    # oracle/__init__.py): the two crossed combinations
    "cross_gandiva_sched_horus": (lambda: frame(70, 68, 1.5), dict(num_switch=2, num_node_p_switch=2, num_gpu_p_node=8,
                                  _scheme="gandiva", _schedule="horus", num_buffer=3)),
    "cross_horus_sched_gandiva": (lambda: frame(70, 69, 1.5), dict(num_switch=1, num_node_p_switch=3, num_gpu_p_node=8,
                                  _scheme="horus", _schedule="gandiva", num_buffer=1)),
    # --scheme yarn (no packing) under the look-ahead / time-slice / credit schedulers
    "yarn_sched_horus": (lambda: frame(90, 70, 1.5, gpu_choices=[1, 2, 4, 8, 12, 16], gpu_probs=[.3, .2, .2, .15, .1, .05]),
                         dict(num_switch=2, num_node_p_switch=3, num_gpu_p_node=8, _scheme="yarn", _schedule="horus", num_buffer=4)),
    "yarn_sched_gandiva": (lambda: frame(80, 71, 2.0, gpu_choices=[1, 2, 4, 8, 12], gpu_probs=[.3, .3, .2, .1, .1]),
                           dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8, _scheme="yarn", _schedule="gandiva", num_buffer=1)),
    "yarn_sched_horusplus": (lambda: frame(80, 72, 1.5), dict(num_switch=1, num_node_p_switch=4, num_gpu_p_node=8,
                             _scheme="yarn", _schedule="horus+", num_queue=3, num_buffer=5)),
    "horusplus_k5": (lambda: frame(120, 67, 2.0, gpu_choices=[1, 2, 4, 8, 16], gpu_probs=[.3, .3, .2, .1, .1]),
                     dict(num_switch=2, num_node_p_switch=3, num_gpu_p_node=8, _scheme="horus+", _schedule="horus+", num_queue=5, num_buffer=4)),
}


def main(which):
    for name in which:
        build, flags = CASES[name]
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        df = build()
        trace = os.path.join(d, "trace.csv")
        df.to_csv(trace, index=False)
        meta = {"flags": dict(flags), "numpy_seed": mg.SEED,
                "reference": "matthewygf/GPUSchedule @ ea0f1474, run_sim.py --scheme %s --schedule %s" % (flags["_scheme"], flags["_schedule"])}
        with open(os.path.join(d, "horus.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        mg.run_reference(trace, flags, d)
        n_rows = sum(1 for _ in open(os.path.join(d, "cluster.csv"))) - 1
        n_jobs = sum(1 for _ in open(os.path.join(d, "job.csv"))) - 1
        print(f"{name}: ticks={n_rows} finished={n_jobs}")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
