#!/usr/bin/env python
"""execute.py -- sweep driver with the reference's interface (/root/reference/execute.py:5-55):
`do_once(scheme, schedule, num_queue, num_buffer)` launches one run_sim.py process with the same
argument list; `main()` walks a list of configurations.  Only the configurations this engine
implements are in the default sweep (the reference's own list is horus/gandiva heavy, which is
out of scope); pass --trace to point at a trace file.
"""
import argparse
import os
import sys
from subprocess import Popen

HERE = os.path.dirname(os.path.abspath(__file__))


def do_once(scheme, schedule, num_queue, num_buffer, trace_file="data/month.csv",
            num_switch=4, num_nodes_p_switch=32, migrate=True, wait=True):
    trace_tag = os.path.splitext(os.path.basename(trace_file))[0]
    log_sub_dir = "thesis_fitted_" + str(num_buffer) + "_nodes_p_s" + str(num_nodes_p_switch) + "_job_" + trace_tag
    log_path = os.path.join(log_sub_dir, f"{scheme}_{schedule}")
    cmd = [sys.executable, os.path.join(HERE, "run_sim.py"),
           "--num_node_p_switch", str(num_nodes_p_switch),
           "--num_switch", str(num_switch),
           "--scheme", scheme,
           "--trace_file", trace_file,
           "--num_queue", str(num_queue),
           "--num_buffer", str(num_buffer),
           "--schedule", schedule,
           "--enable_network_costs", "False",
           "--enable_migration", str(migrate),
           "--log_path", log_path]
    p = Popen(cmd)
    print("process pid %d: " % p.pid)
    if wait:
        try:
            p.wait()
        except KeyboardInterrupt:
            p.kill()
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trace", default=os.path.join("data", "month.csv"))
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--batched", action="store_true",
                    help="run the whole sweep as replicas of one GPU launch (gpuschedule_b200.sweep)")
    a = ap.parse_args()
    if a.batched:
        sys.path.insert(0, HERE)
        from gpuschedule_b200 import sweep
        return sweep.main(["--trace", a.trace, "--schedule", "fifo", "--repeats", str(a.repeats)])
    schemes = ["yarn"]
    schedules = ["fifo"]
    queues = [1]
    buffers = [1]
    for scheme, schedule, queue in zip(schemes, schedules, queues):
        for buff in buffers:
            for _ in range(a.repeats):
                do_once(scheme, schedule, queue, buff, trace_file=a.trace)


if __name__ == "__main__":
    main()
