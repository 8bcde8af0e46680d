#!/usr/bin/env python
"""bench.py -- simulated events/sec on the BASELINE workload.

Workload (config.workload): R independent replicas per GPU of the 100k-job
synthetic trace (SURVEY 8d generator, distinct seeds) on the 4x32x8 cluster,
fifo policy + yarn placement -- the only policy/scheme pair the reference can
execute and therefore the one with a pinned bit-exact oracle.
One "step" = one full simulation of every replica (one persistent-kernel
launch, one warp per replica).  A single replica is latency bound by
construction (<= 1 placement per simulated tick), so throughput comes from
replicas -- the sweeps the reference's execute.py runs serially.

  value  : events/s with traces resident in HBM (device-timed with CUDA events
           around the engine kernel + state reset, max over ranks)
  e2e    : events/s through the C ABI with HOST buffers: every step re-uploads
           every trace (pinned staging -> H2D) and reads back every statistics
           row, job record, finish order and placement span (D2H)
  impl=reference : the CPU oracle port (oracle/gsched_oracle.c, a restatement of
           the reference's Python loop) on all host cores, one replica per thread

Launch:  python bench.py --gpus N --steps K --warmup W   (torchrun for N > 1)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "simulated events/sec (100k-job trace, 4x32x8 cluster)"
UNIT = "events/s"
BASE_SEED = 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def fast_table(n_jobs, seed, rate=0.5):
    """Synthetic columns -> JobTable without the pandas round trip (rows are generated in
    arrival order already; only tie order inside a tick differs from a CSV ingest)."""
    from gpuschedule_b200 import ingest, tracegen
    c = tracegen.synth_columns(n_jobs, seed=seed, rate=rate)
    nt = c["normalized_time"].astype(np.float64) / 10000.0
    return ingest.JobTable(
        n=n_jobs, label=None, num_gpu_text=None,
        arrive_tick=np.ceil(nt).astype(np.int32), submit=nt.astype(np.int32),
        gpus=c["used_gpus"].astype(np.int32), gpu_per_task=c["gpu_per_container"].astype(np.int32),
        duration=np.ascontiguousarray(c["minutes"] * 0.5), mem_bytes=c["memory_max"].astype(np.int64),
        util_avg=c["gpu_utilization_avg"], util_max=c["gpu_utilization_max"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_to_done(eng, rows_cap):
    while True:
        eng.run(0, rows_cap)
        if all(eng.stats(s).done for s in range(eng.nsims)):
            return


def ours(args):
    import torch
    import torch.distributed as dist
    from gpuschedule_b200 import capi
    from gpuschedule_b200.log_manager import JOB_DTYPE, ROW_DTYPE, SPAN_DTYPE

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"             # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier_sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    R, n = args.replicas, args.jobs
    cluster = capi.make_cluster(4, 32, 8)
    from gpuschedule_b200 import policies as gpol

    def policy_for(table):
        if args.policy == "fifo":
            return capi.make_policy("fifo")
        if args.policy == "sjf":
            return capi.make_policy("sjf")
        if args.policy in ("dlas", "dlas-gpu"):
            return capi.make_policy(args.policy, num_queue=4, queue_limit=[3600, 7200, 18000])
        return capi.make_policy("gittins", gittins_delta=3250.0,
                                gittins_table=gpol.build_gittins_table(gpol.gittins_samples(table), 3250.0))
    t0 = time.time()
    from gpuschedule_b200 import dist as gdist
    tables = [fast_table(n, sd) for sd in gdist.replica_seeds(rank, world, R, base=BASE_SEED)]
    log(f"[rank {rank}] generated {R} traces of {n} jobs in {time.time() - t0:.1f}s")
    eng = capi.Engine(device=local, nsims=R)
    eng.set_engine(args.engine)
    if args.span_budget > 0:
        eng.set_span_budget(args.span_budget)
    pols = [policy_for(t) for t in tables]
    for r in range(R):
        eng.config(r, cluster, pols[r])
        eng.load_trace(r, tables[r])

    # ---- warm-up (also sizes the per-replica row window so one launch completes a run)
    run_to_done(eng, 0)
    ticks = [eng.stats(r).ticks for r in range(R)]
    rows_cap = max(ticks) + 64
    for _ in range(max(args.warmup - 1, 2)):
        eng.reset()
        run_to_done(eng, rows_cap)

    # ---- timed: K steps, traces resident in HBM
    sampler = ClockSampler(local)
    launches0 = eng.launch_count()
    barrier_sync()
    sampler.start()
    w0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.reset()
        run_to_done(eng, rows_cap)
        dev_ms += eng.stats(0).kernel_ms
    barrier_sync()
    wall_ms = (time.perf_counter() - w0) * 1e3
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    st = [eng.stats(r) for r in range(R)]
    events_rank = sum(s.events for s in st)
    ticks_rank = sum(s.ticks for s in st)
    evals_rank = sum(s.placement_evals for s in st)
    spans_rank = 0
    for r in range(min(R, 4)):
        spans_rank += len(eng.fetch_spans(r)[1])
    spans_rank = spans_rank / min(R, 4) * R
    if args.policy != "fifo":
        # secondary measurement (event-driven policy kernel): device-timed value only
        dev_ms = max_over_ranks(dev_ms)
        events_all = sum_over_ranks(events_rank)
        if rank == 0:
            import oracle
            c0 = time.perf_counter()
            ref = oracle.run_policy(cluster, pols[0], tables[0])
            t_cpu = time.perf_counter() - c0
            assert ref.events == st[0].events and ref.ticks == st[0].ticks, "engine/oracle disagree"
            print(json.dumps({"metric": METRIC, "value": events_all / (dev_ms / args.steps / 1e3), "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "data": "synthetic", "dtype": "int32/int64",
                              "config": {"workload": f"{n}-job synthetic trace x {R} replicas/GPU, 4x32x8, {args.policy}",
                                         "policy": args.policy, "replicas_per_gpu": R, "jobs_per_replica": n,
                                         "events_per_step": events_all, "rows_per_step": ticks_rank * world},
                              "kernel": "gs_policy_kernel (thread per replica)", "clocks": clocks,
                              "gpu_launches": int(launches),
                              "cpu_baseline": {"value": ref.events / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                                               "sample": "1 replica, oracle/policy_oracle.c (restatement, not reference code)"}}),
                  flush=True)
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    events_all = sum_over_ranks(events_rank)
    value = events_all / (dev_ms / args.steps / 1e3)

    if args.value_only:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": dev_ms / args.steps, "replicas_per_gpu": R}), flush=True)
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (gs_tick_kernel), per launch, this rank's GPU
    # algorithmic bytes: job table in (28 B/job) + job record + finish order out (28 B/job)
    # + one 16-B span per (job,node) + one 64-B statistics row per tick   
    alg_bytes = R * n * 56 + spans_rank * 16 + ticks_rank * 64
    peak, peak_src = peaks()
    ach = alg_bytes / (dev_ms / args.steps / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:      # DRAM bytes per launch from the committed ncu --set full capture (per replica, scaled to this R)
        tj = json.load(open(os.path.join(REPO, "profiles", "tick_kernel_traffic.json")))
        if tj["jobs_per_replica"] == n and args.policy == "fifo" and args.engine in (0, 1):
            traffic = tj["dram_bytes_per_replica"] * R
            traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at %d replicas "
                           "(profiles/tick_kernel_traffic.json), scaled per replica to %d" % (tj["replicas"], R))
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "gs_tick_kernel", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "ticks_per_s": ticks_rank * world / (dev_ms / args.steps / 1e3),
                "candidate_evals_per_s": evals_rank * world / (dev_ms / args.steps / 1e3)}

    # ---- end to end through the C ABI with host buffers: K host threads, each driving its own
    # engine handle (own stream + pinned staging) over a slice of the replicas; every step re-uploads
    # every trace from host memory and reads back every row, record, finish order and span.
    T = max(ticks) + 64
    span_cap = int(max(len(eng.fetch_spans(r)[1]) for r in range(min(R, 8))) * 1.25) + 4096
    eng.close()                                          # free the HBM of the value run first
    packed = [t.packed() for t in tables]                # host-resident 32-byte records (built at ingest time)
    K = max(1, min(args.e2e_threads, R))
    slices = [list(range(k, R, K)) for k in range(K)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    results = [None] * K
    errors = []
    start_evt = threading.Barrier(K + 1)

    def worker(k):
        try:
            mine = slices[k]
            e = capi.Engine(device=local, nsims=len(mine))
            e.set_engine(args.engine)
            if args.span_budget > 0:
                e.set_span_budget(args.span_budget)
            pins = [capi.PinnedBuffer(T * ROW_DTYPE.itemsize), capi.PinnedBuffer(n * JOB_DTYPE.itemsize),
                    capi.PinnedBuffer(n * 4), capi.PinnedBuffer((n + 1) * 8), capi.PinnedBuffer(span_cap * SPAN_DTYPE.itemsize)]
            rows_v, jobs_v = pins[0].view(ROW_DTYPE, T), pins[1].view(JOB_DTYPE, n)
            ord_v, off_v, sp_v = pins[2].view(np.int32, n), pins[3].view(np.int64, n + 1), pins[4].view(SPAN_DTYPE, span_cap)
            for i in range(len(mine)):
                e.config(i, cluster)
            ph = dict(load=0.0, run=0.0, fetch=0.0)
            h2d = d2h = chk = ev = 0
            for step in range(e2e_steps + 1):            # step 0 = untimed warm-up (allocations)
                if step == 1:
                    start_evt.wait()                      # all threads + main: timed region starts
                    ph = dict(load=0.0, run=0.0, fetch=0.0)
                h2d = d2h = ev = 0
                c0 = time.perf_counter()
                for i, r in enumerate(mine):
                    e.load_trace_packed(i, packed[r])
                    h2d += n * 32
                c1 = time.perf_counter(); ph["load"] += c1 - c0
                run_to_done(e, rows_cap)
                c2 = time.perf_counter(); ph["run"] += c2 - c1
                for i in range(len(mine)):
                    s = e.stats(i)
                    rows, recs, order, span_off, spans = e.fetch_all(i, rows_v, jobs_v, ord_v, off_v, sp_v)
                    chk += int(rows["finished"][-1]) + int(recs["end"][0]) + int(order[-1]) + len(spans)
                    d2h += s.ticks * 64 + n * 24 + s.finished * 4 + len(spans) * 16 + (n + 1) * 8
                    ev += s.events
                ph["fetch"] += time.perf_counter() - c2
            results[k] = (ph, h2d, d2h, chk, ev)
            for pb in pins:
                pb.free()
            e.close()
        except Exception as exc:                          # surface worker failures in the main thread
            errors.append(exc)
            try:
                start_evt.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(K)]
    for t in threads:
        t.start()
    barrier_sync()
    start_evt.wait()                                      # released together with the workers' timed steps
    w0 = time.perf_counter()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    barrier_sync()
    e2e_ms = max_over_ranks((time.perf_counter() - w0) * 1e3) / e2e_steps
    h2d = sum(r[1] for r in results); d2h = sum(r[2] for r in results)
    checksum = sum(r[3] for r in results)
    assert sum(r[4] for r in results) == events_rank, "e2e run simulated a different number of events"
    ph = {k: max(r[0][k] for r in results) for k in ("load", "run", "fetch")}
    e2e = {"value": events_all / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(sum_over_ranks(h2d)), "d2h_bytes_per_step": int(sum_over_ranks(d2h)),
           "steps": e2e_steps, "host_threads": K,
           "timing": "wall clock between barrier+synchronize, max over ranks",
           "phase_ms_per_step_slowest_thread": {k: v * 1e3 / e2e_steps for k, v in ph.items()},
           "checksum": checksum}

    # ---- secondary measurements (rank 0, N=1): the event-driven policies of BASELINE configs 1-3 on
    # the same cluster (device-timed, 1 warm-up + 1 timed run each) and the stateless scoring kernel
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        rp = min(R, args.policy_replicas)
        for name, njobs in (("sjf", min(n, 10000)), ("dlas-gpu", n), ("gittins", n)):
            tabs = tables[:rp] if njobs == n else [fast_table(njobs, sd) for sd in gdist.replica_seeds(0, 1, rp, base=BASE_SEED)]
            if name == "gittins":
                tabs = tabs[:max(1, rp // 2)]
            save = args.policy
            args.policy = name
            pl = [policy_for(t) for t in tabs]
            args.policy = save
            with capi.Engine(device=local, nsims=len(tabs)) as pe:
                for i, t in enumerate(tabs):
                    pe.config(i, cluster, pl[i])
                    pe.load_trace_packed(i, t.packed())
                run_to_done(pe, 0)
                cap = max(pe.stats(i).ticks for i in range(len(tabs))) + 64
                pe.reset()
                run_to_done(pe, cap)
                ms = pe.stats(0).kernel_ms
                ev = sum(pe.stats(i).events for i in range(len(tabs)))
            extras[name] = {"value": ev / (ms / 1e3), "unit": UNIT, "ms": ms, "replicas": len(tabs), "jobs_per_replica": njobs,
                            "kernel": "gs_dlas_warp_kernel" if name == "dlas-gpu" else "gs_sortpol_warp_kernel",
                            "parity": "engine == oracle/policy_oracle.c == the reference's loop functions executed under stubs (tests/golden/policy_*)"}
        try:
            import contextlib
            import io
            buf = io.StringIO()
            pa = argparse.Namespace(place_jobs=16 * 1024 * 1024, warmup=2, steps=3)
            with contextlib.redirect_stdout(buf):
                place_mode(pa)
            pj = json.loads(buf.getvalue().strip().splitlines()[-1])
            extras["place_batch"] = {"value": pj["value"], "unit": pj["unit"], "kernel_ms": pj["kernel_ms"], "jobs": pj["jobs"],
                                     "candidate_evals_per_s": pj["candidate_evals_per_s"], "roofline_frac": pj["roofline"]["frac"],
                                     "achieved_gbs": pj["roofline"]["achieved"]}
        except Exception as exc:
            extras["place_batch"] = {"error": str(exc)}
        # widening row f1 (horus / gandiva / horus+ engine): its own process with a time limit, so that nothing it
        # does can cost the main measurement
        try:
            import subprocess
            hp = subprocess.run([sys.executable, os.path.abspath(__file__), "--mode", "horus", "--horus-replicas", "1184"],
                                capture_output=True, text=True, timeout=240)
            line = [ln for ln in hp.stdout.splitlines() if ln.startswith("{")]
            extras["horus"] = json.loads(line[-1]) if line else {"error": (hp.stderr or "no output")[-400:]}
        except Exception as exc:
            extras["horus"] = {"error": repr(exc)}

    # ---- CPU baseline: the oracle port, 1 thread, bounded sample (rank 0, N=1 only)
    cpu = None
    cpu_tight = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        t_cpu, ev_cpu, k = 0.0, 0, 0
        while t_cpu < args.cpu_seconds and k < R:
            c0 = time.perf_counter()
            ref = oracle.run_fifo(cluster, tables[k], rows_cap=T + 64, want_spans=False)
            t_cpu += time.perf_counter() - c0
            ev_cpu += ref.events
            assert ref.ticks == ticks[k] and ref.events == st[k].events, "engine/oracle disagree"
            k += 1
        cpu = {"value": ev_cpu / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{k} replica(s) of the {n}-job trace, full runs, oracle/gsched_oracle.c single thread",
               "host_cores": os.cpu_count()}
        # extra yardstick (not the reference's algorithm): the engine's own O(1)-counter algorithm as
        # tight single-thread C, oracle/tight_cpu.c -- the strongest CPU competitor we could write
        tr = oracle.TightRunner(cluster, tables[0])
        tr.run()
        c0 = time.perf_counter(); reps = 0; ev_t = 0
        while time.perf_counter() - c0 < 2.0:
            tk, ev = tr.run(); reps += 1; ev_t += ev
        assert tk == ticks[0]
        cpu_tight = {"value": ev_t / (time.perf_counter() - c0), "unit": UNIT, "cores": 1,
                     "kind": "tight C restatement of the ENGINE's algorithm (oracle/tight_cpu.c), not reference code",
                     "sample": f"{reps} runs of one {n}-job replica"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 (+f64 durations)",
            "data": "synthetic",
            "config": {"workload": f"{n}-job synthetic trace x {R} replicas/GPU (distinct seeds), 4x32x8 cluster, fifo+yarn",
                       "jobs_per_replica": n, "replicas_per_gpu": R, "cluster": "4x32x8", "policy": "fifo",
                       "scheme": "yarn", "parallelism": f"replicas x{world} GPUs, no data-path collective",
                       "l2": "inputs+outputs per step (%.1f GB/GPU) exceed the 126 MB L2" % (alg_bytes / 1e9),
                       "events_per_step": events_all, "ticks_per_step": ticks_rank * world},
            "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "cpu_tight": cpu_tight, "secondary": extras,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def place_mode(args):
    """Secondary measurement: gs_place_batch, the stateless (job x candidate node) scoring kernel.
    b job requests against one 128-node cluster state; kernel-only time from the library's CUDA
    events; algorithmic bytes = 16 B request in + 8 B (first node, nodes used) out per job."""
    from gpuschedule_b200 import capi
    rng = np.random.default_rng(5)
    m, g = 128, 8
    cluster = capi.make_cluster(4, 32, g)
    nodes = np.zeros(m, dtype=capi.NODE_DTYPE)
    for i in range(m):
        k = int(rng.integers(0, g + 1))
        nodes["busy_mask"][i] = sum(1 << int(d) for d in rng.choice(g, size=k, replace=False))
        nodes["cpu_used"][i] = 12 * k
        nodes["mem_used"][i] = 60 * k
    b = args.place_jobs
    jobs = np.zeros(b, dtype=capi.JOBREQ_DTYPE)
    jobs["gpu_per_task"] = 1
    jobs["gpus"] = rng.choice([1, 2, 4, 8, 16, 32], size=b, p=[.35, .2, .2, .15, .07, .03])
    jobs["mem_bytes"] = rng.integers(512, 16384, size=b).astype(np.int64) << 20
    with capi.Engine(device=0, nsims=1) as eng:
        best = None
        for _ in range(args.warmup + args.steps):
            first, used, _, ms = eng.place_batch(cluster, nodes, jobs)
            best = ms if best is None else min(best, ms)
        placed = int((first >= 0).sum())
    peak, src = peaks()
    gbs = b * 24 / (best / 1e3) / 1e9
    print(json.dumps({"metric": "gs_place_batch jobs scored/s (128-node cluster state)", "value": b / (best / 1e3),
                      "unit": "jobs/s", "kernel_ms": best, "jobs": b, "placeable": placed,
                      "candidate_evals_per_s": b * m / (best / 1e3),
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "algorithmic_bytes_per_launch": b * 24, "peak_source": src,
                                   "kernel": "gs_place_kernel", "traffic": None}}), flush=True)


def horus_mode(args):
    """Secondary measurement (widening row f1): the utilisation-aware engine, `--horus-replicas` independent
    horus simulations (one thread each, gs_horus_kernel) of `--horus-jobs`-job traces on a 2x4x8 cluster.
    Replicas have distinct traces and read one common numpy stream (seed 0).  value = events / kernel time
    (library CUDA events); cpu_baseline = oracle/horus_oracle.c on one host core over a sample of replicas."""
    import oracle
    from gpuschedule_b200 import capi
    R, n = args.horus_replicas, args.horus_jobs
    cluster = capi.make_cluster(num_switch=2, num_node_p_switch=4, num_gpu_p_node=8)
    tables = [fast_table(n, BASE_SEED + 1000 + r, rate=1.0) for r in range(R)]
    np.random.seed(0)
    stream = np.random.standard_normal(args.horus_stream)          # one numpy stream, shared by all replicas
    hp = capi.make_horus_params("horus", "horus", 5)
    with capi.HorusEngine(device=0, nsims=R) as eng:
        for r in range(R):
            eng.config(r, cluster, hp)
            eng.load_trace(r, tables[r])
        by_lanes = {}
        for lanes in ((32, 1) if args.horus_both_mappings else (1,)):   # 1 simulation per warp won round 1 (2.8x)
            eng.set_lanes(lanes)
            eng.load_stream(-1, stream)                    # (re)loading the stream starts the replicas over
            eng.run(rows_cap=args.horus_rows)
            by_lanes[lanes] = float(eng.stats(0).kernel_ms)
        st = [eng.stats(r) for r in range(R)]
        ms = min(by_lanes.values())
        events = sum(int(x.events) for x in st)
        ticks = sum(int(x.ticks) for x in st)
        draws = sum(int(x.draws) for x in st)
        assert all(x.done for x in st)
        rows0, util0, flags0, recs0, order0 = eng.fetch(0)
    ref = oracle.run_horus(cluster, tables[0], scheme="horus", schedule="horus", num_buffer=5, seed=0)
    assert rows0.tobytes() == ref.rows.tobytes() and util0.tobytes() == ref.util.tobytes(), "replica 0 differs from the oracle"
    # horus+ (credit queues + k-means, raw word stream): a few replicas against the oracle, replica by replica
    plus = None
    try:
        k = 6
        np.random.seed(1)
        words = np.random.randint(0, 2 ** 32, size=args.horus_words, dtype=np.uint32)
        pp = capi.make_horus_params("horus+", "horus+", 15, 3)
        with capi.HorusEngine(device=0, nsims=k) as eng:
            for r in range(k):
                eng.config(r, cluster, pp)
                eng.load_trace(r, tables[r])
            eng.load_words(-1, words)
            eng.run(rows_cap=args.horus_rows)
            plus_ms = float(eng.stats(0).kernel_ms)
            same = 0
            for r in range(k):
                prow, putil, pflag, precs, porder = eng.fetch(r)
                pref = oracle.run_horus(cluster, tables[r], scheme="horus+", schedule="horus+", num_buffer=15, num_queue=3, seed=1)
                same += int(prow.tobytes() == pref.rows.tobytes() and putil.tobytes() == pref.util.tobytes()
                            and precs.tobytes() == pref.recs.tobytes() and np.array_equal(porder, pref.finish_order))
        plus = {"replicas": k, "identical_to_oracle": same, "kernel_ms": plus_ms}
    except Exception as exc:                                # noqa: BLE001 - reported, never fatal for the horus line
        plus = {"error": repr(exc)}
    sample = min(R, 64)
    t0 = time.perf_counter()
    cpu_ev = sum(oracle.run_horus(cluster, tables[r], scheme="horus", schedule="horus", num_buffer=5, seed=0).events for r in range(sample))
    cpu_s = time.perf_counter() - t0
    # the warp-cooperative mapping (one simulation per warp, all lanes score together): LAST, in its own handle, so
    # that whatever it does cannot touch the numbers above
    coop = None
    try:
        with capi.HorusEngine(device=0, nsims=R) as eng:
            eng.set_lanes(0)
            for r in range(R):
                eng.config(r, cluster, hp)
                eng.load_trace(r, tables[r])
            eng.load_stream(-1, stream)
            eng.run(rows_cap=args.horus_rows)
            cst = [eng.stats(r) for r in range(R)]
            crow, cutil, cflag, crecs, corder = eng.fetch(0)
            coop = {"kernel_ms": float(cst[0].kernel_ms), "events": sum(int(x.events) for x in cst),
                    "events_per_s": sum(int(x.events) for x in cst) / (float(cst[0].kernel_ms) / 1e3),
                    "replica0_identical_to_oracle": bool(crow.tobytes() == ref.rows.tobytes() and cutil.tobytes() == ref.util.tobytes()
                                                         and crecs.tobytes() == ref.recs.tobytes()),
                    "same_event_total_as_scalar_mapping": sum(int(x.events) for x in cst) == events}
    except Exception as exc:                                # noqa: BLE001
        coop = {"error": repr(exc)}
    if isinstance(plus, dict) and "error" not in plus and "error" not in coop:
        try:                                                # horus+ under the cooperative mapping (indexed word stream)
            k = plus["replicas"]
            with capi.HorusEngine(device=0, nsims=k) as eng:
                eng.set_lanes(0)
                for r in range(k):
                    eng.config(r, cluster, pp)
                    eng.load_trace(r, tables[r])
                eng.load_words(-1, words)
                eng.run(rows_cap=args.horus_rows)
                same = 0
                for r in range(k):
                    prow, putil, pflag, precs, porder = eng.fetch(r)
                    pref = oracle.run_horus(cluster, tables[r], scheme="horus+", schedule="horus+", num_buffer=15, num_queue=3, seed=1)
                    same += int(prow.tobytes() == pref.rows.tobytes() and putil.tobytes() == pref.util.tobytes() and precs.tobytes() == pref.recs.tobytes())
                coop["horus_plus"] = {"replicas": k, "identical_to_oracle": same, "kernel_ms": float(eng.stats(0).kernel_ms)}
        except Exception as exc:                            # noqa: BLE001
            coop["horus_plus"] = {"error": repr(exc)}
    print(json.dumps({"metric": "horus simulated events/s (replica batch)", "value": events / (ms / 1e3), "unit": UNIT,
                      "kernel_ms": ms, "replicas": R, "jobs_per_replica": n, "ticks": ticks, "samples_drawn": draws,
                      "samples_per_s": draws / (ms / 1e3), "kernel": "gs_horus_kernel (one simulation per thread)", "kernel_ms_by_lanes_per_warp": by_lanes,
                      "parity": "replica 0 == oracle/horus_oracle.c == reference (tests/golden/horus_*)",
                      "horus_plus_device_check": plus, "cooperative_warp_mapping": coop,
                      "cpu_baseline": {"value": cpu_ev / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                                       "sample": f"{sample} of the {R} replicas, oracle/horus_oracle.c, one thread"}}), flush=True)


def reference(args):
    """The reference arm: the CPU port of the reference's loop on all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import concurrent.futures as cf
    import oracle
    from gpuschedule_b200 import capi
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.jobs
    cluster = capi.make_cluster(4, 32, 8)
    oracle.lib()
    threads = max(1, min(cores, args.cpu_threads or cores))
    if not args.cpu_threads and cores > 4:
        # "all the host threads it can use": container CPU quotas can make fewer threads faster than
        # one per visible core, so pick the count with the best throughput on a short probe
        probe_t = fast_table(min(n, 20000), BASE_SEED)
        probe_cap = int(probe_t.arrive_tick[-1]) + 2 * int(np.ceil(probe_t.duration.max())) + 4096
        best = (0.0, 1)
        k = 1
        while k <= cores:
            with cf.ThreadPoolExecutor(k) as ex:
                t0 = time.perf_counter()
                ev = sum(ex.map(lambda _: oracle.run_fifo(cluster, probe_t, rows_cap=probe_cap, want_spans=False).events, range(k)))
                rate = ev / (time.perf_counter() - t0)
            if rate > best[0]:
                best = (rate, k)
            k *= 2
        threads = best[1]
    tables = [fast_table(n, BASE_SEED + r) for r in range(threads)]

    caps = [int(t.arrive_tick[-1]) + 2 * int(np.ceil(t.duration.max())) + 4096 for t in tables]

    def one(it):
        t, cap = it
        return oracle.run_fifo(cluster, t, rows_cap=cap, want_spans=False).events   # ctypes releases the GIL

    with cf.ThreadPoolExecutor(threads) as ex:
        work = list(zip(tables, caps))
        for _ in range(args.warmup):
            list(ex.map(one, work))
        t0 = time.perf_counter()
        events = 0
        for _ in range(args.steps):
            events += sum(ex.map(one, work))
        dt = time.perf_counter() - t0
    value = events / dt
    runners = [oracle.TightRunner(cluster, t) for t in tables]
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda r: r.run(), runners))
        t0 = time.perf_counter(); ev_t = 0; reps = 0
        while time.perf_counter() - t0 < 3.0:
            ev_t += sum(e for _, e in ex.map(lambda r: r.run(), runners)); reps += 1
        tight_value = ev_t / (time.perf_counter() - t0)
    sample = f"{threads} replicas of the {n}-job trace per step (one per thread), full runs"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
           "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int32/int64 (+f64 durations)", "data": "synthetic",
           "config": {"workload": f"{n}-job synthetic trace, 4x32x8 cluster, fifo+yarn", "jobs_per_replica": n,
                      "cluster": "4x32x8", "policy": "fifo", "scheme": "yarn"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                            "note": "oracle/gsched_oracle.c: C restatement of the reference's Python loop "
                                    "(the Python reference itself: 186 events/s at N=10k, BASELINE.md)"},
           "cpu_tight": {"value": tight_value, "unit": UNIT, "cores": threads,
                         "kind": "tight C restatement of the ENGINE's algorithm (oracle/tight_cpu.c), not reference code",
                         "sample": f"{reps} rounds of {threads} replicas, one per thread"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--jobs", type=int, default=100000)
    ap.add_argument("--replicas", type=int, default=3552, help="replicas per GPU (one warp each)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-threads", type=int, default=16, help="host threads (one engine handle each) in the e2e run")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--value-only", action="store_true", help="kernel experiments: print the device-timed value and stop")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary policy / place_batch measurements")
    ap.add_argument("--policy-replicas", type=int, default=1024)
    ap.add_argument("--span-budget", type=float, default=0.0,
                    help="span-pool records per job (0 = worst case); the trace uses ~1.13")
    ap.add_argument("--policy", default="fifo", choices=["fifo", "sjf", "dlas", "dlas-gpu", "gittins"],
                    help="fifo = the headline (pinned) workload; others = secondary, event-driven policy kernel")
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 warp per replica, 2 lane per replica")
    ap.add_argument("--mode", default="sim", choices=["sim", "place", "horus"], help="place = gs_place_batch micro-benchmark; horus = utilisation-aware engine")
    ap.add_argument("--horus-replicas", type=int, default=2368)
    ap.add_argument("--horus-jobs", type=int, default=60)
    ap.add_argument("--horus-stream", type=int, default=4000000, help="standard-normal samples loaded per replica")
    ap.add_argument("--horus-rows", type=int, default=8192)
    ap.add_argument("--horus-both-mappings", action="store_true", help="also time 32 simulations per warp")
    ap.add_argument("--horus-words", type=int, default=6 << 20, help="raw generator words for the horus+ device check")
    ap.add_argument("--place-jobs", type=int, default=64 * 1024 * 1024)
    args = ap.parse_args()
    if args.mode == "place":
        place_mode(args)
    elif args.mode == "horus":
        horus_mode(args)
    elif args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
