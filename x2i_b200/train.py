"""The attention-distillation train step of X2I (``train/train_qwenvl.py:559-651`` student side, ``:717-816`` teacher
side) on the sm_100a kernels, data-parallel over all ranks.

Per step and per rank, on ITS shard of the batch:
  1. teacher pass (no grad): frozen FLUX conditioned on the T5 / CLIP embeddings, hooks capture every ``blk.attn`` output;
  2. student pass: projector(MLLM hidden states) -> the SAME frozen FLUX (teacher and student are one checkpoint,
     ``train_qwenvl.py:417,:667``) in saving mode, hooks capture the same tensors with autograd history;
  3. KD loss = sum over the 19+19+38 hooked layers of KL(student || teacher) (``:601-620``), backward through all 57 blocks
     (flux_train.py) and the projector (proj._ProjFn);
  4. ONE collective: all-reduce (mean) of the projector gradients (what the reference's DDP does, ``:483``);
  5. clip_grad_norm_(1.0), optimizer step, LR scheduler step (``:625-632``).

The reference's 6+2 teacher/student rank split and its gather/scatter of 1.7 GB/sample of hook tensors
(core/pipeline/train_and_infer.py) are not reproduced: both passes run on every rank, hook tensors never leave the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import dist as xdist
from . import kd


def prepare_latent_image_ids(height: int, width: int, device, dtype):
    """train_qwenvl.py:216-227 (height/width in packed-latent units x2, i.e. 128 for 1024 px)."""
    ids = torch.zeros(height // 2, width // 2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height // 2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width // 2)[None, :]
    return ids.reshape(-1, 3).to(device=device, dtype=dtype)


_IDS_CACHE: Dict = {}


def _cached_ids(S: int, height: int, width: int, device, dtype):
    """Stable (txt_ids, img_ids) tensors per shape: the transformer's RoPE cache keys on their addresses, so rebuilding them every
    step would cost a table rebuild or a host-synchronising content compare per step."""
    key = (S, height, width, str(device), dtype)
    if key not in _IDS_CACHE:
        _IDS_CACHE[key] = (torch.zeros(S, 3, device=device, dtype=dtype), prepare_latent_image_ids(height, width, device, dtype))
    return _IDS_CACHE[key]


def _run_hooked(transformer, lists, **kw):
    handles = []
    lists.append([]); lists.append([]); lists.append([])

    def two(_m, _i, output):
        lists[0].append(output[0]); lists[1].append(output[1])

    def one(_m, _i, output):
        lists[2].append(output)

    for blk in transformer.transformer_blocks:
        handles.append(blk.attn.register_forward_hook(two))
    for blk in transformer.single_transformer_blocks:
        handles.append(blk.attn.register_forward_hook(one))
    try:
        out = transformer(return_dict=False, **kw)[0]
    finally:
        for h in handles:
            h.remove()
    return out


def get_max_numbered_filename(directory: str):
    """train_qwenvl.py:199-203: the largest integer found in the entry names of `directory` (None if there is none)."""
    import os
    import re
    if not os.path.isdir(directory):
        return None
    pattern = re.compile(r"\d+")
    numbers = [int(pattern.search(f).group()) for f in os.listdir(directory) if pattern.search(f)]
    return max(numbers) if numbers else None


def resume_projector(proj, output_dir: str):
    """The reference's auto-resume (train_qwenvl.py:404-410, :534-536): load
    ``{output_dir}/{max step}/diffusion_pytorch_model.bin`` into the projector (weights only: the reference saves neither optimizer,
    LR-scheduler nor RNG state) and return that step as the new ``global_step``; None when there is no checkpoint."""
    import os
    from .proj import load_projector_state
    step = get_max_numbered_filename(output_dir)
    if step is None:
        return None
    path = os.path.join(output_dir, str(step), "diffusion_pytorch_model.bin")
    load_projector_state(proj, torch.load(path, map_location="cpu"))
    return step


def distill_step(proj, transformer, batch: Dict[str, torch.Tensor], optimizer=None, lr_scheduler=None, max_grad_norm: float = 1.0,
                 temperature: float = 3.0, guidance_scale: float = 3.5, height: int = 128, width: int = 128, group=None,
                 stacked: bool = False, micro_step: int = 0, gradient_accumulation_steps: int = 1, bucket=None, timings=None):
    """One distillation (micro-)step on this rank's batch shard.

    Gradient accumulation follows train_qwenvl.py:561,:625-632: every micro-step adds its gradients to ``.grad`` (the loss is NOT
    divided by the number of micro-steps), and the optimizer fires when ``micro_step % gradient_accumulation_steps == 0``.  The
    reference's DDP all-reduces on every micro-step (no ``no_sync``); the mean over ranks is linear, so here the ONE all-reduce
    runs on the accumulated gradients right before the clip -- same update, 1/gas of the traffic.
    bucket: an x2i_b200.dist.GradBucket over the projector parameters (gradients live in one flat buffer: all-reduce and clip
    without flatten/scatter copies); timings: optional list receiving the (start, end) CUDA events of the all-reduce.

    batch: ``latents`` [B, L_img, 64] (pure noise, t = 1.0 in the reference), ``timestep`` [B] (x1000 scale, as the
    reference's scheduler emits; divided by 1000 here like ``:580``), ``text_embeddings`` [B, C, S, H] (all-layer MLLM
    hidden states, student conditioning), ``prompt_embeds_t5`` [B, S, 4096] and ``pooled_clip`` [B, 768] (teacher
    conditioning).  Returns the detached loss tensor (no host sync).
    stacked=True reproduces the reference's torch.stack of the hook lists before the loss (API parity check)."""
    dev = batch["latents"].device
    dt = transformer.dtype
    B = batch["latents"].shape[0]
    S = batch["prompt_embeds_t5"].shape[1]
    txt_ids, img_ids = _cached_ids(S, height, width, dev, dt)
    guidance = torch.full((B,), guidance_scale, device=dev, dtype=dt) if transformer.config.guidance_embeds else None
    common = dict(hidden_states=batch["latents"].to(dt), timestep=batch["timestep"] / 1000, txt_ids=txt_ids, img_ids=img_ids,
                  guidance=guidance)
    kd_teacher, kd_student = [], []
    from .ops import nvtx
    with torch.no_grad(), nvtx("x2i.distill.teacher"):
        _run_hooked(transformer, kd_teacher, encoder_hidden_states=batch["prompt_embeds_t5"].to(dt),
                    pooled_projections=batch["pooled_clip"].to(dt), **common)
    with nvtx("x2i.distill.projector"):
        add_text_embeds, prompt_embeds = proj(batch["text_embeddings"])
    with nvtx("x2i.distill.student"):
        _run_hooked(transformer, kd_student, encoder_hidden_states=prompt_embeds.to(dt), pooled_projections=add_text_embeds.to(dt),
                    **common)
    with nvtx("x2i.distill.kd_loss"):
        if stacked:
            loss = kd.attention_distillation_loss(kd_teacher, kd_student, temperature, verbose=False)
        else:
            t_all = kd_teacher[0] + kd_teacher[1] + kd_teacher[2]
            s_all = kd_student[0] + kd_student[1] + kd_student[2]
            loss, _valid = kd.kd_loss_layers(t_all, s_all, temperature)
    with nvtx("x2i.distill.backward"):
        loss.backward()
    if micro_step % max(1, gradient_accumulation_steps) != 0:
        return loss.detach()  # accumulate only (train_qwenvl.py:561)
    params = [p for p in proj.parameters() if p.requires_grad]
    with nvtx("x2i.distill.allreduce"):
        if bucket is not None:
            bucket.allreduce_mean_(group=group, timings=timings)
        else:
            xdist.allreduce_mean_grads_(params, group=group)
    if optimizer is not None:
        if max_grad_norm is not None:
            if bucket is not None:
                bucket.clip_grad_norm_(max_grad_norm)
            else:
                torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
        optimizer.step()
        if lr_scheduler is not None:
            lr_scheduler.step()
        if bucket is not None:
            bucket.zero_()
        else:
            optimizer.zero_grad(set_to_none=True)
    return loss.detach()


def save_projector_checkpoint(proj, output_dir: str, global_step: int) -> str:
    """The reference's on-disk contract (train_qwenvl.py:641-647): {output_dir}/{step}/diffusion_pytorch_model.bin."""
    import os
    path = os.path.join(output_dir, f"{global_step}")
    os.makedirs(path, exist_ok=True)
    fn = os.path.join(path, "diffusion_pytorch_model.bin")
    torch.save({k: v.detach().cpu() for k, v in proj.state_dict().items()}, fn)
    return fn


def synthetic_batch(B: int, device, cfg, C: int = 37, S: int = 512, H: int = 2048, L_img: int = 4096, seed: int = 0):
    """Synthetic inputs of BASELINE config 4: N(0,1) embeddings, noise latents, t = 1000."""
    g = torch.Generator(device=device).manual_seed(seed)
    bf = torch.bfloat16
    r = lambda *s: torch.randn(*s, device=device, generator=g).to(bf)  # noqa: E731
    return dict(latents=r(B, L_img, cfg["in_channels"]), timestep=torch.full((B,), 1000.0, device=device, dtype=bf),
                text_embeddings=r(B, C, S, H), prompt_embeds_t5=r(B, S, cfg["joint_attention_dim"]),
                pooled_clip=r(B, cfg["pooled_projection_dim"]))
