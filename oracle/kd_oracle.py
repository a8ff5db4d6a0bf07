"""ORACLE (test infrastructure only) for the attention-distillation loss.

Restates ``/root/reference/train/train_qwenvl.py``:
  * ``normalize`` :58-61  -- (x - mean) / (1e-7 + unbiased std) over the last dim;
  * the loss loop :601-620 -- per layer ``F.kl_div(softmax(norm(teacher)/T).log(),
    softmax(norm(student)/T), reduction='batchmean')`` i.e. KL(student || teacher) summed
    over every element and divided by the batch size; a layer whose term is inf/nan is
    skipped; the 19+19+38 layer terms are added up.
``normalize`` is PINNED against the reference function itself (imported with stubbed
third-party modules by oracle/make_golden.py -> tests/golden/kd_normalize.pt); the loop is
inline in ``train()`` and cannot be called, so it is restated from the source lines above.
"""
import torch
import torch.nn.functional as F


def normalize(logit: torch.Tensor) -> torch.Tensor:
    mean = logit.mean(dim=-1, keepdim=True)
    std = logit.std(dim=-1, keepdim=True)  # unbiased (N-1)
    return (logit - mean) / (1e-7 + std)


def kd_layer_term(teacher: torch.Tensor, student: torch.Tensor, temperature: float = 3.0) -> torch.Tensor:
    """One layer's term; teacher/student are [B, L, D]."""
    log_pt = F.softmax(normalize(teacher) / temperature, dim=-1).log()
    ps = F.softmax(normalize(student) / temperature, dim=-1)
    return F.kl_div(log_pt, ps, reduction="batchmean")


def kd_loss(teacher_layers, student_layers, temperature: float = 3.0):
    """teacher_layers / student_layers: sequences of [B, L, D] tensors (any L per layer).
    Returns (loss, list_of_skipped_layer_indices)."""
    loss = 0
    skipped = []
    for i, (t, s) in enumerate(zip(teacher_layers, student_layers)):
        term = kd_layer_term(t, s, temperature)
        if torch.isinf(term).any() or torch.isnan(term).any():
            skipped.append(i)
        else:
            loss = loss + term
    return loss, skipped


def kd_loss_stacked(kd_t0, kd_t1, kd_t2, kd_s0, kd_s1, kd_s2, temperature: float = 3.0):
    """The reference's exact call shape: stacked [B, n_layers, L, D] tensors for the double-block
    image stream (0), text stream (1) and the single blocks (2); order of accumulation as in
    train_qwenvl.py:603-620 (img_i, txt_i interleaved, then singles)."""
    loss = 0
    for i in range(kd_t0.shape[1]):
        for t, s in ((kd_t0, kd_s0), (kd_t1, kd_s1)):
            term = kd_layer_term(t[:, i], s[:, i], temperature)
            if not (torch.isinf(term).any() or torch.isnan(term).any()):
                loss = loss + term
    for i in range(kd_t2.shape[1]):
        term = kd_layer_term(kd_t2[:, i], kd_s2[:, i], temperature)
        if not (torch.isinf(term).any() or torch.isnan(term).any()):
            loss = loss + term
    return loss
