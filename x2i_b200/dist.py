"""Multi-GPU plumbing for the X2I hot path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Inference shards the image batch across ranks with NO data-path collective ("replicas only": every rank holds the
full frozen FLUX weights).  Distillation training is plain data parallelism with ONE collective per optimiser step: the
all-reduce of the projector gradients (the reference wraps only the projector in DDP, train/train_qwenvl.py:483).
The reference's teacher/student rank split and its gather/scatter of hook tensors (core/pipeline/train_and_infer.py:
31-122, 1.7 GB per sample) are deliberately not reproduced: teacher and student are the same frozen checkpoint, so both
run on every rank (multi-GPU).
"""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def dist_info() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched directly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = dist_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of a global batch for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_mean_grads_(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """The single exchange step of distillation training: grads <- mean over ranks, flattened into one bucket so it is
    one all-reduce (57-69 MB bf16 for the X2I projectors).  Reproduces DDP's semantics for the reference's loss
    normalisation: every rank's loss is sum/bsz_local ('batchmean') and DDP averages over ranks.  Returns #elements."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
