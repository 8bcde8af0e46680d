#!/usr/bin/env python
"""run_sim.py -- same command line as the reference simulator
(/root/reference/run_sim.py:19-94,1710-1757): same flags, same ./log/<log_path>/<timestamp>/
output directory with cluster.csv, job.csv, cpu.csv, gpu.csv, memory.csv, network.csv and
output.log -- but Scheduler.start() runs on the GPU (gpuschedule_b200/libgsched.so).

  python run_sim.py --num_switch 4 --num_node_p_switch 32 --num_gpu_p_node 8 \
        --scheme yarn --schedule fifo --trace_file trace.csv --log_path out
"""
import datetime
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gpuschedule_b200 import flags  # noqa: E402

FLAGS = flags.define_simulator_flags()
flags.DEFINE_integer("seed", -1, "numpy seed for the avg_gpu_utilization column (-1: unseeded, like the reference)")


def main(log_manager):
    from gpuschedule_b200 import infrastructure as cluster
    from gpuschedule_b200 import jobs as jobs_mod
    from gpuschedule_b200 import schedule as sche

    from gpuschedule_b200 import capi
    capi.warm_device_async(FLAGS.device)                   # CUDA context creation overlaps the trace ingest
    infrastructure = cluster.Infrastructure(FLAGS)
    log_manager.init(infrastructure)
    jq_manager = jobs_mod.JobQueueManager(FLAGS, FLAGS.trace_file)
    jobs_manager = jobs_mod.JobsManager(FLAGS, jq_manager)
    scheduler = sche.Scheduler(infrastructure, jobs_manager, log_manager,
                               enable_migration=FLAGS.enable_migration)
    if FLAGS.seed >= 0:
        import numpy
        numpy.random.seed(FLAGS.seed)
    stats = scheduler.start()
    logging.info("ticks=%d events=%d finished=%d kernel_ms=%.3f", stats.ticks, stats.events,
                 stats.finished, stats.kernel_ms)
    return stats


if __name__ == "__main__":
    from gpuschedule_b200 import log_manager as lm
    logging.basicConfig(format="%(asctime)s p%(process)s {%(module)s:%(lineno)d} %(levelname)s: %(message)s",
                        level=logging.INFO)
    execution_id = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S-%f")
    output_dir = os.path.join("log", FLAGS.log_path, execution_id)
    os.makedirs(output_dir, exist_ok=True)
    handler = logging.FileHandler(filename=os.path.join(output_dir, "output.log"), mode="w")
    handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
    logging.getLogger().addHandler(handler)
    main(lm.LogManager(output_dir, FLAGS))
    logging.getLogger().removeHandler(handler)
    sys.exit(0)
