#!/usr/bin/env python
"""execute.py -- sweep driver with the reference's interface (/root/reference/execute.py:5-55):
`do_once(scheme, schedule, num_queue, num_buffer)` launches one run_sim.py process with the same
argument list; `main()` walks the reference's own list of configurations (execute.py:47-55: horus+ with 3 / 4 / 5
queues, horus, gandiva, yarn+fifo; every look-ahead width of its `buffers` list; 3 repeats).  --batched runs the same
list as replicas of two engine launches instead of 108 processes; pass --trace to point at a trace file.
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
    if schedule == "horus+":
        log_path = os.path.join(log_path, "k" + str(num_queue))
    options = dict(num_node_p_switch=num_nodes_p_switch, num_switch=num_switch, scheme=scheme, trace_file=trace_file,
                   num_queue=num_queue, num_buffer=num_buffer, schedule=schedule, enable_network_costs=False,
                   enable_migration=migrate, log_path=log_path)          # the argument list of execute.py:19-33
    cmd = [sys.executable, os.path.join(HERE, "run_sim.py")]
    for name, value in options.items():
        cmd += ["--" + name, str(value)]
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
    schemes = ["horus+", "horus+", "horus+", "horus", "gandiva", "yarn"]          # execute.py:48-51
    queues = [3, 4, 5, 1, 1, 1]
    schedules = ["horus+", "horus+", "horus+", "horus", "gandiva", "fifo"]
    buffers = [15, 15, 15, 1, 1, 1]
    if a.batched:
        sys.path.insert(0, HERE)
        from gpuschedule_b200 import sweep
        tag = os.path.splitext(os.path.basename(a.trace))[0]
        sets = []
        for scheme, schedule, queue in zip(schemes, schedules, queues):
            for buff in buffers:
                for _ in range(a.repeats):
                    sub = os.path.join("thesis_fitted_" + str(buff) + "_nodes_p_s32_job_" + tag, f"{scheme}_{schedule}")
                    if schedule == "horus+":
                        sub = os.path.join(sub, "k" + str(queue))
                    sets.append(sweep.make_flags(trace_file=a.trace, scheme=scheme, schedule=schedule, num_queue=queue, num_buffer=buff,
                                                 num_switch=4, num_node_p_switch=32, enable_migration=True, log_path=sub))
        for out_dir, st in sweep.run_batched(sets):
            print(f"{out_dir}: ticks={st.ticks} events={st.events} finished={st.finished}")
        return
    for scheme, schedule, queue in zip(schemes, schedules, queues):
        for buff in buffers:
            for _ in range(a.repeats):
                do_once(scheme, schedule, queue, buff, trace_file=a.trace)


if __name__ == "__main__":
    main()
