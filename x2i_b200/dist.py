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


class GradBucket:
    """All gradients of a (small) trainable module as views of ONE flat buffer -- the shape DDP's ``gradient_as_bucket_view`` gives
    the reference's projector (train/train_qwenvl.py:483).  autograd accumulates straight into the views (in place, which is
    also what gradient accumulation over micro-steps needs, :561/:625), so the single exchange step of a train step is one
    ``all_reduce`` on the buffer itself: no flatten copy before it and no scatter copies after it, and the clip norm is one
    reduction over the same buffer.  ``timings`` (optional list) receives (start, end) CUDA events around the collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        p0 = self.params[0]
        if any(p.dtype != p0.dtype or p.device != p0.device for p in self.params):
            raise ValueError("GradBucket: parameters must share dtype and device")
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=p0.dtype, device=p0.device)
        self.attach_()

    def attach_(self):
        """(Re-)point every .grad at its slice of the buffer (needed again after optimizer.zero_grad(set_to_none=True))."""
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        return self

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero_(self):
        self.flat.zero_()
        if self.params[0].grad is None or self.params[0].grad.data_ptr() != self.flat.data_ptr():
            self.attach_()

    def allreduce_mean_(self, group=None, timings=None):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return self
        ev = None
        if timings is not None and self.flat.is_cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)  # mean over ranks inside the collective
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
        if ev is not None:
            ev[1].record()
            timings.append(ev)
        return self

    def clip_grad_norm_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_ semantics (2-norm over all gradients, scale by max_norm / (norm + 1e-6) when above) as
        one reduction + one scale over the flat buffer, without a host sync."""
        total = torch.linalg.vector_norm(self.flat.float(), 2.0)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(coef.to(self.flat.dtype))
        return total


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
