"""World-size-2 gloo test of the multi-GPU host logic (replica sharding + reductions)."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from gpuschedule_b200 import dist as gd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = gd.Reducer(world)
    seeds = gd.replica_seeds(rank, world, 5, base=1)
    events = 3 * 1000 * len(seeds)                 # every replica finishes 1000 jobs
    ms = 10.0 + 7.0 * rank                         # rank 1 is slower
    red.barrier()
    total = red.sum(events)
    worst = red.max(ms)
    lo, hi = gd.shard_range(11, rank, world)
    # sharded single simulation: every rank ends up with every rank's exchange-buffer handle, in rank order
    handles = gd.exchange_comm_handles(bytes([rank + 1]) * 64, world)
    assert handles == [bytes([r + 1]) * 64 for r in range(world)]
    out.put((rank, seeds, total, worst, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, t0, w0, sh0), (r1, s1, t1, w1, sh1) = res
    assert s0 == [1, 2, 3, 4, 5] and s1 == [6, 7, 8, 9, 10]          # disjoint, contiguous
    assert t0 == t1 == 3 * 1000 * 10                                   # whole-job aggregate on every rank
    assert w0 == w1 == 17.0                                            # max over ranks
    assert sh0 == (0, 6) and sh1 == (6, 11)


def test_reference_arm_runs_on_rank0_only(monkeypatch, capsys):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    args = type("A", (), dict(jobs=200, steps=1, warmup=0, cpu_threads=1, replicas=4736, config="c1"))()
    bench.reference(args)                          # non-zero ranks exit without work or output
    assert capsys.readouterr().out == ""
    monkeypatch.setenv("RANK", "0")
    bench.reference(args)
    import json
    line = json.loads(capsys.readouterr().out)
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["config"] == bench.config_block(200, 4736)         # the same `config` the GPU arm prints


def test_replica_seed_validation():
    from gpuschedule_b200 import dist as gd
    with pytest.raises(ValueError):
        gd.replica_seeds(2, 2, 4)
    assert gd.shard_range(10, 0, 3) == (0, 4) and gd.shard_range(10, 2, 3) == (7, 10)


def test_chunk_owner_partitions_the_runnable_list():
    from gpuschedule_b200 import dist as gd
    for world in (1, 2, 4, 8):
        owners = [gd.chunk_owner(c, world) for c in range(64)]
        assert set(owners) == set(range(world))
        for r in range(world):
            assert owners.count(r) == 64 // world          # balanced round robin
    with pytest.raises(ValueError):
        gd.exchange_comm_handles(b"short", 2)
