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


def _arena_worker(rank, world, port, q):
    """Training exchange: the LoRA-gradient arena of a small UNet, bucketed sum all-reduce in reverse layer order."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oracle.configs import UNET_CONFIGS
    from t2v_turbo_b200 import dist as d
    from t2v_turbo_b200.lora_train import arena_for_unet
    from t2v_turbo_b200.unet import UNetModel
    d.init_replicas("gloo")
    with torch.device("meta"):
        unet = UNetModel(**UNET_CONFIGS["small"]["cfg"])
    arena = arena_for_unet(unet, "cpu", r=64)
    g = torch.Generator().manual_seed(100 + rank)
    arena.grads.copy_(torch.randn(arena.padded, generator=g))
    mine = arena.grads.clone()
    red = d.ArenaReducer(arena.grads, n_buckets=5)
    # the backward completes layers from the last to the first: gradients become final from the end of the arena
    for i in reversed(range(0, len(arena.shapes), 7)):
        red.ready(arena.offsets[i])
    nbytes = red.finish()
    q.put((rank, arena.numel, nbytes, mine.numpy(), arena.grads.numpy().copy()))   # numpy: pickled by value
    d.barrier()
    d.shutdown()


def test_lora_grad_arena_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, n0, b0, mine0, red0), (_, n1, b1, mine1, red1) = [(a, b, c, torch.from_numpy(d), torch.from_numpy(e)) for a, b, c, d, e in res]
    assert n0 == n1 and b0 == b1 == red0.numel() * 4           # every byte of the arena crossed exactly once
    assert torch.equal(red0, red1)                               # identical on both ranks
    torch.testing.assert_close(red0, mine0 + mine1, rtol=0, atol=0)


def test_arena_census_vc2():
    """575 LoRA target layers, 117 142 176 trainable values at r = 64 (468.6 MB of fp32 gradients per step)."""
    from oracle.configs import VC2_UNET
    from t2v_turbo_b200.lora_train import arena_for_unet
    from t2v_turbo_b200.unet import UNetModel
    with torch.device("meta"):
        unet = UNetModel(**VC2_UNET)
        arena = arena_for_unet(unet, "meta", r=64)
    assert len(arena.shapes) == 2 * 575 and arena.numel == 117_142_176
