"""Multi-GPU plumbing.

Inference: REPLICAS ONLY.

Each video (prompt, seed) is independent — no cross-sample op exists in the UNet, scheduler or VAE — so N
GPUs run N independent pipeline replicas, one process per GPU, and the data path has no collective.  The only
communication is for measurement: a barrier around the timed region and a MAX-reduce of the per-rank device
times (the slowest replica defines the whole-job time).  Backend: NCCL on GPUs, gloo in the CPU tests.

Training (consistency distillation, train_t2v_turbo_v1_lora.py:862,1190): pure data parallel with ONE exchange — the
LoRA gradients.  All 575 layers' weight gradients live in one contiguous fp32 arena (lora_train.LoraArena: 117 142 176
values = 468.6 MB for the VC2 UNet at r = 64), so the DDP mean is one sum all-reduce over that buffer (NCCL over
NVLink / NVSwitch), issued in a few large buckets in reverse layer order as the backward produces them so that the
transfer overlaps the remaining backward GEMMs; the 1 / world factor is folded into the fused AdamW launch.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_replicas(backend: str | None = None, device: torch.device | None = None) -> tuple[int, int]:
    """Initialise the process group from the torchrun environment (no-op for a single process)."""
    rank, _, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if (device is not None and device.type == "cuda") else "gloo"
        kw = {"device_id": device} if backend == "nccl" and device is not None else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier(device: torch.device | None = None) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(values, device: torch.device | None = None) -> list:
    """Element-wise maximum over ranks of a list of floats (per-rank device times in ms)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


class ArenaReducer:
    """Bucketed sum all-reduce of a flat gradient arena, overlapped with the backward that fills it.

    The arena is cut into `n_buckets` contiguous ranges.  The backward walks the layers in reverse, so gradients complete
    from the END of the arena towards its start: `ready(offset)` tells the reducer that every gradient at or above `offset`
    is final, and each bucket that lies entirely above it is all-reduced asynchronously (torch.distributed launches NCCL on
    its own stream after the work already enqueued on the current stream).  `finish()` reduces what is left and waits.
    The result is the SUM over ranks; divide by `world` in the optimizer (LoraArena.adamw_step(grad_scale=1 / world))."""

    def __init__(self, grads: torch.Tensor, n_buckets: int = 8):
        assert grads.dim() == 1 and grads.is_contiguous()
        self.grads = grads
        n = grads.numel()
        step = -(-n // max(1, n_buckets))
        step = (step + 1023) // 1024 * 1024                      # 4 KB aligned bucket boundaries
        self.bounds = [(a, min(a + step, n)) for a in range(0, n, step)]
        self.next = len(self.bounds) - 1                          # buckets are reduced from the last to the first
        self.works = []
        self.bytes = 0

    @property
    def active(self):
        return dist.is_initialized() and dist.get_world_size() > 1

    def _reduce(self, i):
        a, b = self.bounds[i]
        self.bytes += (b - a) * self.grads.element_size()
        if self.active:
            self.works.append(dist.all_reduce(self.grads[a:b], op=dist.ReduceOp.SUM, async_op=True))

    def ready(self, offset: int):
        while self.next >= 0 and self.bounds[self.next][0] >= offset:
            self._reduce(self.next)
            self.next -= 1

    def finish(self):
        self.ready(0)
        for w in self.works:
            w.wait()
        self.works = []
        self.next = len(self.bounds) - 1
        n, self.bytes = self.bytes, 0
        return n


def shard_prompts(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of a list of independent prompts for this rank (ragged tail goes to the low ranks)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
