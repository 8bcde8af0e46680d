"""Replica-batched sweeps: what the reference's execute.py runs as N sequential processes
(/root/reference/execute.py:47-55) becomes ONE engine handle with N replicas advanced by the same
kernel launches.  Every replica keeps its own flags, trace, policy and output directory, and
writes the same files a single run_sim.py invocation would."""
from __future__ import annotations

import argparse
import datetime
import os
import types

import numpy as np

from . import capi, rngcol
from .infrastructure import Infrastructure
from .jobs import JobQueueManager, JobsManager
from .log_manager import LogManager
from .schedule import Scheduler

DEFAULTS = dict(trace_file="tf_job.csv", log_path="batched", scheme="yarn", schedule="fifo", pack=False,
                num_switch=1, num_node_p_switch=32, enable_network_costs=False, enable_migration=False,
                bandwidth=1250, internode_latency=0.015, gpu_memory_capacity=32, num_queue=1, num_buffer=5,
                num_gpu_p_node=8, num_cpu_p_node=128, mem_p_node=512, cluster_spec=None, device=0,
                queue_limit="3600,7200,18000", gittins_delta=3250.0, seed=-1)


def make_flags(**overrides):
    """A flags namespace with run_sim.py's defaults (run_sim.py:19-94) plus overrides."""
    unknown = set(overrides) - set(DEFAULTS)
    if unknown:
        raise ValueError(f"unknown flags: {sorted(unknown)}")
    return types.SimpleNamespace(**{**DEFAULTS, **overrides})


def _is_utilisation_aware(fl):
    return fl.scheme in Scheduler.UTILISATION_AWARE or fl.schedule in Scheduler.UTILISATION_AWARE


def run_batched_horus(flag_sets, device=0, out_root="log", chunk=1 << 21, rows_cap=1 << 16):
    """The horus / horus+ / gandiva configurations of a sweep as replicas of ONE gs_horus launch.  Every replica
    owns a numpy RandomState (seeded with flags.seed when >= 0, fresh entropy otherwise -- the reference's repeats
    are unseeded) whose stream it consumes exactly like the single-run path does with numpy's global one, so a
    seeded replica writes the same bytes as `run_sim.py --seed s`.  Returns [(output_dir, stats)]."""
    sims = []
    for fl in flag_sets:
        infra = Infrastructure(fl)
        jm = JobsManager(fl, JobQueueManager(fl, fl.trace_file))
        rs = np.random.RandomState(fl.seed if getattr(fl, "seed", -1) >= 0 else None)
        raw = fl.schedule == "horus+"
        draw = (lambda k, rs=rs: rs.randint(0, 2 ** 32, size=k, dtype=np.uint32)) if raw else (lambda k, rs=rs: rs.standard_normal(k))
        sims.append(dict(fl=fl, infra=infra, jm=jm, raw=raw, draw=draw, stream=draw(chunk), rows_cap=rows_cap))
    results = []
    with capi.HorusEngine(device=device, nsims=len(sims)) as eng:
        for i, sm in enumerate(sims):
            fl = sm["fl"]
            eng.config(i, sm["infra"].gs_cluster(), capi.make_horus_params(fl.scheme, fl.schedule, int(fl.num_buffer), int(fl.num_queue)))
            eng.load_trace(i, sm["jm"].table)
            (eng.load_words if sm["raw"] else eng.load_stream)(i, sm["stream"])
        cap = rows_cap
        while True:
            for sm in sims:
                sm["ran_cap"] = cap                       # the row window this launch really gives every replica
            try:
                eng.run(rows_cap=cap)
                break
            except capi.GsError as e:
                if e.code != capi.GS_ERR_CAPACITY:
                    raise
                for i, sm in enumerate(sims):             # finished replicas keep their results; the short ones start over
                    st = eng.stats(i)
                    if st.status != capi.GS_ERR_CAPACITY:
                        continue
                    if st.ticks < sm["ran_cap"]:          # ran out of samples: continue this replica's stream
                        sm["stream"] = np.concatenate([sm["stream"], sm["draw"](len(sm["stream"]))])
                    else:                                 # ran out of rows
                        sm["rows_cap"] = 2 * sm["ran_cap"]
                    (eng.load_words if sm["raw"] else eng.load_stream)(i, sm["stream"])
                cap = max(sm["rows_cap"] for sm in sims)
        for i, sm in enumerate(sims):
            fl, infra, jm = sm["fl"], sm["infra"], sm["jm"]
            rows, util, flags_arr, recs, order = eng.fetch(i)
            stats = eng.stats(i)
            stamp = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S-%f")
            out_dir = os.path.join(out_root, fl.log_path, f"{stamp}-r{i}")
            os.makedirs(out_dir, exist_ok=True)
            lm = LogManager(out_dir, fl)
            lm.init(infra)
            cl = infra.gs_cluster()
            m, g = cl.num_switch * cl.num_node_p_switch, cl.num_gpu_p_node
            lm.write_cluster_rows(rows, rngcol.sampled_utilization_text(util, flags_arr), m * g * cl.gpu_mem_cap_mib)
            lm.write_horus_job_rows(jm.table, recs, order)
            results.append((out_dir, stats))
    return results


def run_batched(flag_sets, device=0, out_root="log"):
    """Run every configuration of `flag_sets` (list of flags namespaces) as one replica each.
    Returns [(output_dir, stats)] in the order of `flag_sets`; the utilisation-aware configurations go through
    run_batched_horus (their own engine handle), the others through one gs_run batch."""
    aware = [i for i, fl in enumerate(flag_sets) if _is_utilisation_aware(fl)]
    if aware:
        plain = [i for i in range(len(flag_sets)) if i not in set(aware)]
        out = [None] * len(flag_sets)
        for idx, res in zip(aware, run_batched_horus([flag_sets[i] for i in aware], device, out_root)):
            out[idx] = res
        if plain:
            for idx, res in zip(plain, run_batched([flag_sets[i] for i in plain], device, out_root)):
                out[idx] = res
        return out
    sims = []
    for fl in flag_sets:
        infra = Infrastructure(fl)
        jm = JobsManager(fl, JobQueueManager(fl, fl.trace_file))
        sched = Scheduler(infra, jm, None)
        sims.append((fl, infra, jm, sched.make_policy(jm.table)))
    results = []
    with capi.Engine(device=device, nsims=len(sims)) as eng:
        for i, (fl, infra, jm, pol) in enumerate(sims):
            eng.config(i, infra.gs_cluster(), pol)
            eng.load_trace(i, jm.table)
        rows_all = eng.run_all()
        for i, (fl, infra, jm, pol) in enumerate(sims):
            recs, order = eng.fetch_jobs(i)
            span_off, spans = eng.fetch_spans(i)
            stats = eng.stats(i)
            stamp = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S-%f")
            out_dir = os.path.join(out_root, fl.log_path, f"{stamp}-r{i}")
            os.makedirs(out_dir, exist_ok=True)
            lm = LogManager(out_dir, fl)
            lm.init(infra)
            cl = infra.gs_cluster()
            m, g = cl.num_switch * cl.num_node_p_switch, cl.num_gpu_p_node
            if getattr(fl, "seed", -1) >= 0:
                np.random.seed(fl.seed)
            util = rngcol.utilization_text(len(rows_all[i]), m, g, jm.table, recs, span_off, spans)
            lm.write_cluster_rows(rows_all[i], util, m * g * cl.gpu_mem_cap_mib)
            lm.write_job_rows(jm.table, recs, order)
            results.append((out_dir, stats))
    return results


def main(argv=None):
    ap = argparse.ArgumentParser(description="run a sweep of simulator configurations as one batched GPU launch")
    ap.add_argument("--trace", nargs="+", required=True, help="trace CSV file(s)")
    ap.add_argument("--schedule", nargs="+", default=["fifo"])
    ap.add_argument("--scheme", default=None, help="placement scheme; default: the schedule's own (yarn for fifo / sjf / dlas / gittins)")
    ap.add_argument("--num_buffer", type=int, default=5)
    ap.add_argument("--num_switch", type=int, default=4)
    ap.add_argument("--num_node_p_switch", type=int, default=32)
    ap.add_argument("--num_queue", type=int, default=4)
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--seed", type=int, default=-1)
    a = ap.parse_args(argv)
    sets = []
    for tr in a.trace:
        for sc in a.schedule:
            for rep in range(a.repeats):
                tag = os.path.splitext(os.path.basename(tr))[0]
                scheme = a.scheme or (sc if sc in Scheduler.UTILISATION_AWARE else "yarn")
                sets.append(make_flags(trace_file=tr, schedule=sc, scheme=scheme, num_switch=a.num_switch,
                                       num_node_p_switch=a.num_node_p_switch, num_queue=a.num_queue, num_buffer=a.num_buffer,
                                       log_path=os.path.join(f"batched_{tag}", f"{scheme}_{sc}"),
                                       seed=a.seed if a.seed < 0 else a.seed + rep))
    for out_dir, st in run_batched(sets):
        print(f"{out_dir}: ticks={st.ticks} events={st.events} finished={st.finished}")


if __name__ == "__main__":
    main()
