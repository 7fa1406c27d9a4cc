"""Multi-GPU plumbing for the inference path: REPLICAS ONLY.

Each video (prompt, seed) is independent — no cross-sample op exists in the UNet, scheduler or VAE — so N
GPUs run N independent pipeline replicas, one process per GPU, and the data path has no collective.  The only
communication is for measurement: a barrier around the timed region and a MAX-reduce of the per-rank device
times (the slowest replica defines the whole-job time).  Backend: NCCL on GPUs, gloo in the CPU tests.
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


def shard_prompts(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of a list of independent prompts for this rank (ragged tail goes to the low ranks)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
