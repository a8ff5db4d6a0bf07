"""Attention-distillation loss of X2I (``train/train_qwenvl.py:58-61`` normalize, ``:186-214`` hooks, ``:601-620`` loss)
as ONE fused row-wise sm_100a kernel per direction, exposed as a differentiable PyTorch function.

    loss = sum_layers [ KL( softmax(norm(student_l)/T) || softmax(norm(teacher_l)/T) ) summed over elements / B ]
    (layers whose term is inf/nan are skipped, exactly like the reference's guard)

The reference evaluates ~15 eager kernels per layer on stacked copies of the hook tensors; here the row statistics,
both softmaxes and the KL are computed in registers in one pass over teacher and student (fwd: 2 reads; bwd: 2 reads +
1 write), with a deterministic two-stage reduction.  Gradients flow to the student only (the teacher is frozen).
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops
from ._lib import X2IError


def cast_hook_list(unet, lists):
    """Register the reference's forward hooks (train_qwenvl.py:206-214): lists[0]/[1] collect the (img, txt) outputs of
    every double block's ``attn``, lists[2] the output of every single block's ``attn``."""
    lists.append([]); lists.append([]); lists.append([])

    def two(model, input, output):
        lists[0].append(output[0])
        lists[1].append(output[1])

    def one(model, input, output):
        lists[2].append(output)

    for net in unet.transformer_blocks:
        net.attn.register_forward_hook(two)
    for net in unet.single_transformer_blocks:
        net.attn.register_forward_hook(one)


def _segments(shape, device):
    """Stacked [B, n_layers, L, D] -> B*n_layers segments of L rows; segment (b, i) belongs to layer i."""
    B, n, L, _ = shape
    starts = torch.arange(0, B * n + 1, device=device, dtype=torch.int64) * L
    layer = torch.arange(n, device=device, dtype=torch.int32).repeat(B)
    return starts, layer, L


class _KDStacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, temperature):
        B, n, L, D = teacher.shape
        t2 = teacher.detach().to(torch.bfloat16).contiguous().view(-1, D)
        s2 = student.detach().to(torch.bfloat16).contiguous().view(-1, D)
        starts, layer, max_rows = _segments(teacher.shape, teacher.device)
        loss, terms, valid = ops.kd_loss_fwd(t2, s2, starts, layer, n, B, temperature)
        ctx.save_for_backward(t2, s2, starts, layer, valid)
        ctx.meta = (B, max_rows, temperature, student.shape, student.dtype)
        ctx.mark_non_differentiable(terms, valid)
        return loss, terms, valid

    @staticmethod
    def backward(ctx, dloss, _dterms, _dvalid):
        t2, s2, starts, layer, valid = ctx.saved_tensors
        B, max_rows, temperature, shape, dtype = ctx.meta
        g = ops.kd_loss_bwd(t2, s2, starts, layer, max_rows, B, valid, dloss.float().contiguous(), temperature)
        return None, g.view(shape).to(dtype), None


def kd_loss_stacked(teacher: torch.Tensor, student: torch.Tensor, temperature: float = 3.0):
    """One stacked hook tensor pair [B, n_layers, L, D] -> (loss, layer_terms[n_layers], valid[n_layers])."""
    if teacher.shape != student.shape or teacher.dim() != 4:
        raise X2IError("kd_loss_stacked: teacher and student must be [B, n_layers, L, D] of equal shape")
    return _KDStacked.apply(teacher, student, float(temperature))


def attention_distillation_loss(KD_teacher: Sequence, KD_student: Sequence, temperature: float = 3.0, verbose: bool = True):
    """The reference's loss over its three hook groups (train_qwenvl.py:590-620).

    KD_teacher / KD_student: three entries (double-block image stream, double-block text stream, single blocks), each
    either the stacked tensor [B, n_layers, L, D] (``torch.stack(hook_list, dim=1)``) or the hook list itself.
    Prints ``down_feature{,1,2}:{i}`` for skipped layers like the reference."""
    total = 0
    for gi, (t, s) in enumerate(zip(KD_teacher, KD_student)):
        if isinstance(t, (list, tuple)):
            t = torch.stack(list(t), dim=1)
        if isinstance(s, (list, tuple)):
            s = torch.stack(list(s), dim=1)
        loss, _terms, valid = kd_loss_stacked(t, s, temperature)
        total = total + loss
        if verbose:
            bad = (valid == 0).nonzero().flatten().tolist()  # host sync only when asked to report
            for i in bad:
                print(f"down_feature{['', '1', '2'][gi]}:{i}")
    return total
