"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: replicas only — prompt sharding, barrier,
max-over-ranks timing.  The data path itself has no collective to test."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from t2v_turbo_b200 import dist as d
    r, w = d.init_replicas("gloo")
    assert (r, w) == (rank, world)
    d.barrier()
    times = d.max_over_ranks([10.0 + rank, 5.0 - rank])       # rank-dependent "device times"
    shard = list(d.shard_prompts(5, rank, world))
    d.barrier()
    d.shutdown()
    q.put((rank, times, shard))


def test_replicas_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == [11.0, 5.0]               # max over ranks, identical on every rank
    assert res[0][2] == [0, 1, 2] and res[1][2] == [3, 4]      # disjoint, exhaustive shards


def test_single_process_is_noop():
    from t2v_turbo_b200 import dist as d
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    assert d.init_replicas() == (0, 1)
    assert d.max_over_ranks([3.5]) == [3.5]
    assert list(d.shard_prompts(3, 0, 1)) == [0, 1, 2]
