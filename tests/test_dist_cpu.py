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


def _v2_worker(rank, world, port, q):
    """v2 full fine-tune step, data parallel (SURVEY §8 f1 + e): each rank runs train_step_v2 on ITS OWN batch (through the CPU
    restatements of the kernels, tests/mock_ops.py) with dist.ArenaReducer hooked into the backward; returns its local gradient
    (captured before the exchange), the reduced arena and the stepped parameters."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from _pytest.monkeypatch import MonkeyPatch
    import mock_ops
    mp_ = MonkeyPatch()
    mock_ops.install(mp_)
    from oracle.configs import UNET_CONFIGS
    from oracle.weights import seeded_state_dict
    from t2v_turbo_b200 import dist as d
    from t2v_turbo_b200.distill_v2 import V2Step, train_step_v2
    from t2v_turbo_b200.full_train import FullUNet
    from t2v_turbo_b200.scheduler import T2VTurboScheduler
    from t2v_turbo_b200.unet import UNetModel
    torch.set_num_threads(2)
    d.init_replicas("gloo")
    spec = UNET_CONFIGS["small_motion"]
    m = UNetModel(**spec["cfg"])
    m.load_state_dict(seeded_state_dict(m.state_dict(), spec["weight_seed"]), strict=True)
    s = FullUNet(m.eval()).eval()
    s.pack()
    step = V2Step(s, T2VTurboScheduler(linear_start=0.00085, linear_end=0.012))
    g = torch.Generator().manual_seed(500 + rank)
    shape = (1,) + tuple(spec["x_shape"][1:])
    batch = dict(index=torch.tensor([120 + 40 * rank]), prompt_emb=torch.randn(1, spec["ctx_len"], spec["cfg"]["context_dim"], generator=g),
                 **{k: torch.randn(shape, generator=g) for k in ("z_t", "cond_teacher_out", "uncond_teacher_out", "score")})
    red = d.ArenaReducer(s.arena.grads, n_buckets=6)
    local = {}
    orig_ready = red.ready

    def ready(off):                # keep this rank's own gradient of every range at the moment it is handed to the exchange
        hi = local.get("lo", s.arena.padded)
        local.setdefault("parts", []).append((off, s.arena.grads[off:hi].clone()))
        local["lo"] = off
        orig_ready(off)
    red.ready = ready
    p0 = s.arena.params.clone()
    train_step_v2(step, batch, lr=1e-4, temporal_lr_scale=2.0, reducer=red, world=world, max_grad_norm=None, fixed=dict(w=torch.tensor([7.0 + rank])))
    mine = torch.zeros_like(s.arena.grads)
    for off, part in local["parts"]:
        mine[off:off + part.numel()] = part
    q.put((rank, mine.numpy(), s.arena.grads.numpy().copy(), (s.arena.params - p0).numpy(), [tuple(r) for r in s.arena.runs]))
    d.barrier()
    d.shutdown()
    mp_.undo()


def test_v2_full_finetune_step_data_parallel_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_v2_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, mine0, red0, dp0, runs), (_, mine1, red1, dp1, _) = [(a, torch.from_numpy(b), torch.from_numpy(c), torch.from_numpy(e), f) for a, b, c, e, f in res]
    assert mine0.norm() > 0 and mine1.norm() > 0 and not torch.allclose(mine0, mine1)      # different batches, different gradients
    assert torch.equal(red0, red1)                                                          # the exchange: identical sums on both ranks
    torch.testing.assert_close(red0, mine0 + mine1, rtol=0, atol=0)
    assert torch.equal(dp0, dp1) and dp0.abs().max() > 0                                    # replicas stay in lock-step after AdamW
    # AdamW's first step moves a weight by ~lr * sign(mean gradient): lr 1e-4 in the other group, 2e-4 in the temporal group
    mean_g = (mine0 + mine1) / 2
    for lo, hi, temporal in runs:
        sel = mean_g[lo:hi].abs() > 1e-6
        if sel.any():
            expect = -(2e-4 if temporal else 1e-4) * torch.sign(mean_g[lo:hi][sel])
            torch.testing.assert_close(dp0[lo:hi][sel], expect, rtol=2e-2, atol=1e-7)


def test_full_arena_census_vc2():
    """the v2 student trains all 1 413 653 060 parameters of the motion-conditioned VC2 UNet (1 413 366 340 + motion_cond_proj
    + combine_proj): 5.65 GB of fp32 gradients per step, laid out in backward-completion order, 33 optimizer-group runs."""
    from oracle.configs import VC2_UNET
    from t2v_turbo_b200.full_train import FullArena
    from t2v_turbo_b200.unet import UNetModel
    with torch.device("meta"):
        unet = UNetModel(**{**VC2_UNET, "motion_cond_proj_dim": 256})
        arena = FullArena(unet, "meta")
    assert arena.numel == 1_413_653_060 and len(arena.names) == 1487 and arena.padded - arena.numel < 4 * len(arena.names)
    assert sum(arena.is_temporal) == 416 and len(arena.runs) == 33 and arena.padded * 4 / 1e9 > 5.65
    assert arena.first_offset("out") > arena.first_offset("output_blocks.11") > arena.first_offset("output_blocks.0") > arena.first_offset("middle_block") \
        > arena.first_offset("init_attn") > arena.first_offset("input_blocks.11") > arena.first_offset("input_blocks.0") > 0
