"""`Scheduler` -- drop-in for core.scheduling.schedule.Scheduler
(/root/reference/core/scheduling/schedule.py:12-215): same constructor, same
`.start()`.  The whole tick loop runs on the GPU behind the C ABI; the host only
formats the result through LogManager (and replays the RNG column)."""
from __future__ import annotations

import logging
import time

from . import capi, policies, rngcol


class Scheduler:
    def __init__(self, infrastructure, jobs_manager, log_manager, enable_migration=False):
        self.infrastructure = infrastructure
        self.jobs_manager = jobs_manager
        self.log_manager = log_manager
        self.placement = infrastructure.flags.scheme
        self.schedule = infrastructure.flags.schedule
        self.enable_migration = enable_migration
        self.stats = None

    def make_policy(self, table):
        """gs_policy from the flags (run_sim.py:37-49 names; README.md:57-62 thresholds)."""
        flags = self.infrastructure.flags
        kw = {}
        if self.schedule in ("dlas", "dlas-gpu"):
            limits = [float(x) for x in str(getattr(flags, "queue_limit", "3600,7200,18000")).split(",") if x]
            kw = dict(num_queue=max(int(getattr(flags, "num_queue", 1)), 1), queue_limit=limits)
        elif self.schedule == "gittins":
            delta = float(getattr(flags, "gittins_delta", 3250.0))
            kw = dict(gittins_delta=delta,
                      gittins_table=policies.build_gittins_table(policies.gittins_samples(table), delta))
        return capi.make_policy(self.schedule, self.placement if self.placement in capi.SCHEMES else "yarn", **kw)

    UTILISATION_AWARE = ("horus", "horus+", "gandiva")

    def _start_utilisation_aware(self):
        """--scheme horus|gandiva: score-based placement with look-ahead / time slicing (include/gsched_horus.h).
        The reference samples numpy's global stream inside these decisions; the engine gets the same stream as
        standard-normal values drawn HERE from numpy's global generator, so `numpy.random.seed(s)` before
        start() reproduces a seeded reference run bit for bit (and an unseeded run is random, as there)."""
        import numpy as np
        t0 = time.time()
        infra, table = self.infrastructure, self.jobs_manager.table
        flags = infra.flags
        cluster = infra.gs_cluster()
        params = capi.make_horus_params(self.placement, self.schedule, int(getattr(flags, "num_buffer", 5)),
                                        int(getattr(flags, "num_queue", 1)))
        # horus+ also draws integers (k-means seeding), so it needs the raw generator words; the others take the
        # cheaper standard-normal form.  Either way the chunks continue numpy's GLOBAL stream.
        raw = self.schedule == "horus+"
        draw = (lambda k: np.random.randint(0, 2 ** 32, size=k, dtype=np.uint32)) if raw else np.random.standard_normal
        stream = draw(1 << 21)
        rows_cap = 1 << 16
        with capi.HorusEngine(device=getattr(flags, "device", 0), nsims=1) as eng:
            eng.config(0, cluster, params)
            eng.load_trace(0, table)
            while True:
                (eng.load_words if raw else eng.load_stream)(0, stream)
                try:
                    eng.run(rows_cap=rows_cap)
                    break
                except capi.GsError as e:
                    if e.code != capi.GS_ERR_CAPACITY:
                        raise
                    if eng.stats(0).ticks >= rows_cap:            # ran out of rows
                        rows_cap *= 2
                    else:                                         # ran out of samples: continue numpy's stream
                        stream = np.concatenate([stream, draw(len(stream))])
            rows, util, flags_arr, recs, order = eng.fetch(0)
            self.stats = eng.stats(0)
        m = cluster.num_switch * cluster.num_node_p_switch
        g = cluster.num_gpu_p_node
        self.log_manager.write_cluster_rows(rows, rngcol.sampled_utilization_text(util, flags_arr), m * g * cluster.gpu_mem_cap_mib)
        logging.info("Total Time Taken in seconds: %d" % (time.time() - t0))
        self.log_manager.write_horus_job_rows(table, recs, order)
        self.rows, self.recs, self.finish_order = rows, recs, order
        return self.stats

    def start(self):
        # horus / horus+ / gandiva as scheme (placement routine) or as schedule (scheduler + score function): the
        # utilisation-aware engine; --scheme yarn with such a schedule runs ms_yarn_placement under that scheduler
        if self.placement in self.UTILISATION_AWARE or self.schedule in self.UTILISATION_AWARE:
            return self._start_utilisation_aware()
        t0 = time.time()
        infra, table = self.infrastructure, self.jobs_manager.table
        cluster = infra.gs_cluster()
        policy = self.make_policy(table)
        with capi.Engine(device=getattr(infra.flags, "device", 0), nsims=1) as eng:
            eng.config(0, cluster, policy)
            eng.load_trace(0, table)
            rows = eng.run_all()[0]
            recs, order = eng.fetch_jobs(0)
            span_off, spans = eng.fetch_spans(0)
            self.stats = eng.stats(0)
        m = cluster.num_switch * cluster.num_node_p_switch
        g = cluster.num_gpu_p_node
        # the sampled column spends most of its time inside numpy's generator and the library (both release the GIL):
        # the text of the other columns and of job.csv is prepared meanwhile
        from concurrent.futures import ThreadPoolExecutor
        from . import log_manager as lm
        cap = m * g * cluster.gpu_mem_cap_mib
        with ThreadPoolExecutor(1) as pool:
            static = pool.submit(lm.cluster_static_columns, rows, cap)
            job_text = pool.submit(lm.job_lines, table, recs, order) if len(order) else None
            util = rngcol.utilization_text(len(rows), m, g, table, recs, span_off, spans)
            self.log_manager.write_cluster_rows(rows, util, cap, static.result())
        logging.info("Total Time Taken in seconds: %d" % (time.time() - t0))
        self.log_manager.write_job_rows(table, recs, order, job_text.result() if job_text else None)
        self.rows, self.recs, self.finish_order = rows, recs, order
        return self.stats
