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


class _KDLayers(torch.autograd.Function):
    """Sum over layers of the reference's per-layer KL term, straight from the hook LISTS (no torch.stack copies of the
    1.6 GB/sample hook tensors, train_qwenvl.py:590-592): tensors = (teacher_0..n-1, student_0..n-1), each [B, L_i, D]."""

    @staticmethod
    def forward(ctx, temperature, n, *tensors):
        teachers, students = tensors[:n], tensors[n:]
        dev = teachers[0].device
        total = torch.zeros((), device=dev, dtype=torch.float32)
        saved, metas = [], []
        valids = []
        for t, s in zip(teachers, students):
            B, L, D = t.shape
            t2 = t.detach().to(torch.bfloat16).contiguous().view(-1, D)
            s2 = s.detach().to(torch.bfloat16).contiguous().view(-1, D)
            starts = torch.arange(0, B + 1, device=dev, dtype=torch.int64) * L
            layer = torch.zeros(B, device=dev, dtype=torch.int32)
            loss, _terms, valid = ops.kd_loss_fwd(t2, s2, starts, layer, 1, B, temperature)
            total = total + loss
            saved += [t2, s2, starts, layer, valid]
            metas.append((B, L, s.shape, s.dtype))
            valids.append(valid)
        ctx.save_for_backward(*saved)
        ctx.meta = (temperature, n, metas)
        valid_all = torch.cat(valids)
        ctx.mark_non_differentiable(valid_all)
        return total, valid_all

    @staticmethod
    def backward(ctx, dloss, _dvalid):
        temperature, n, metas = ctx.meta
        sv = ctx.saved_tensors
        dl = dloss.float().contiguous()
        grads = []
        for i, (B, L, shape, dtype) in enumerate(metas):
            t2, s2, starts, layer, valid = sv[5 * i:5 * i + 5]
            g = ops.kd_loss_bwd(t2, s2, starts, layer, L, B, valid, dl, temperature)
            grads.append(g.view(shape).to(dtype))
        return (None, None) + (None,) * n + tuple(grads)


def kd_loss_layers(teacher_layers: Sequence[torch.Tensor], student_layers: Sequence[torch.Tensor], temperature: float = 3.0):
    """(loss, valid[n_layers]) over per-layer hook tensors [B, L_i, D]; same value as the stacked form (each layer's term is
    sum / B, inf/nan layers skipped) without stacking."""
    if len(teacher_layers) != len(student_layers) or not teacher_layers:
        raise X2IError("kd_loss_layers: need equally long, non-empty teacher / student lists")
    return _KDLayers.apply(float(temperature), len(teacher_layers), *teacher_layers, *student_layers)
