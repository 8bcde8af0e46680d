"""One gittins simulation on two GPUs (include/gsched.h: gs_comm_*; BASELINE config C4): rank r evaluates the index for
its chunks of the runnable list and stores the results into the peer's buffer over NVLink inside the persistent
kernel.  Every rank must end with exactly the bytes of the single-GPU run -- and of the oracle.  Needs two GPUs
(gpurun --gpus 2); skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, n_jobs):
    try:
        import torch
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import oracle
        from gpuschedule_b200 import capi, ingest, tracegen
        from gpuschedule_b200 import dist as gd
        from gpuschedule_b200 import policies as gpol
        cluster = capi.make_cluster(4, 32, 8)
        table = ingest.table_from_columns(tracegen.synth_columns(n_jobs, seed=77, rate=1.2))
        pol = capi.make_policy("gittins", gittins_delta=3250.0,
                               gittins_table=gpol.build_gittins_table(gpol.gittins_samples(table), 3250.0))

        def run(eng):
            eng.config(0, cluster, pol)
            eng.load_trace(0, table)
            rows = eng.run_all()[0]
            recs, order = eng.fetch_jobs(0)
            return rows.tobytes(), recs.tobytes(), order.tobytes(), int(eng.stats(0).events)

        with capi.Engine(device=rank, nsims=1) as e1:
            single = run(e1)
        with capi.Engine(device=rank, nsims=1) as e2:
            handles = gd.exchange_comm_handles(e2.comm_prepare(table.n), world, torch.device("cuda", rank))
            e2.comm_init(rank, handles)
            e2.comm_set_min_runnable(0)                  # exchange on EVERY event, however short the runnable list
            sharded = run(e2)
            exchanges, us = e2.comm_stats()
            e2.reset()                                   # a second run continues the exchange counter
            again = run(e2)
            e2.comm_set_min_runnable(48)                 # adaptive: only events with more than 48 runnable jobs exchange
            e2.reset()
            adaptive = run(e2)
            ex_adaptive, _ = e2.comm_stats()
        ref = oracle.run_policy(cluster, pol, table)
        ok = (sharded == single and again == single and adaptive == single and single[0] == ref.rows.tobytes()
              and single[1] == ref.recs.tobytes() and 0 < ex_adaptive < exchanges)
        dist.barrier()
        dist.destroy_process_group()
        out.put((rank, ok, exchanges, us, single[3]))
    except Exception as exc:                              # noqa: BLE001 -- reported to the parent
        out.put((rank, False, -1, repr(exc), 0))


def test_sharded_gittins_is_bit_identical_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, 4000)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, ok, exchanges, us, events in res:
        assert ok, (rank, us)
        assert exchanges > 1000 and events > 0          # one exchange per event
    assert res[0][2] == res[1][2]                        # both ranks counted the same exchanges
