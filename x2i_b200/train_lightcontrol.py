"""The LightControl train step of X2I (``lightcontrol/train_lightcontrol.py:672-775``) on the sm_100a kernels, plain data
parallelism instead of the reference's DeepSpeed ZeRO-2 offload.

Per step and per rank, on ITS shard of the batch:
  1. (no grad) ``vae.encode(pixel_values).latent_dist.sample()`` -> ``(z - shift) * scale`` (``:676-679``), noise, a logit-normal
     timestep per image (``:690-698``), ``z_t = (1 - sigma) z + sigma noise`` (``:703``), ``_pack_latents`` (``:705-711``);
  2. the frozen FLUX transformer with the 19 trainable ControlNeXt nets injected after its first double blocks
     (``:732-743``; the hint is the style image itself, quirk 9 of SURVEY.md Appendix C) -- ``FluxTransformer2DModel`` routes this
     through ``flux_train.FluxTrainFn`` with the control tokens as differentiable inputs;
  3. flow-matching loss ``mean((v_pred - (noise - z))^2)`` over the unpacked prediction (``:745-763``, weighting "none");
  4. backward: hand-written transformer backward -> control-token gradients -> ControlNeXt backward (conv dgrad / wgrad,
     GroupNorm backward kernels);
  5. ONE collective: all-reduce (mean) of the control-net gradients (19 x 3.6 M parameters), ``clip_grad_norm_``, optimizer step
     (``:769-775``).
The MLLM + projector that produce ``prompt_embeds`` / ``pooled_prompt_embeds`` (``:713-722``) are outside this step: the batch
carries them, exactly what ``proj_t5(text_embeddings)`` returns.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import dist as xdist
from ._lib import X2IError
from .pipeline import FluxPipeline


def compute_density_for_timestep_sampling(weighting_scheme: str, batch_size: int, logit_mean: float = 0.0, logit_std: float = 1.0,
                                          mode_scale: float = 1.29, generator=None, device="cpu"):
    """diffusers.training_utils.compute_density_for_timestep_sampling [D031] (train_lightcontrol.py:690-696)."""
    if weighting_scheme == "logit_normal":
        u = torch.randn(batch_size, generator=generator, device=device) * logit_std + logit_mean
        return torch.sigmoid(u)
    u = torch.rand(batch_size, generator=generator, device=device)
    if weighting_scheme == "mode":
        u = 1 - u - mode_scale * (torch.cos(math.pi * u / 2) ** 2 - 1 + u)
    return u


def train_sigmas(num_train_timesteps: int = 1000, shift: float = 3.0) -> torch.Tensor:
    """The sigma table of FlowMatchEulerDiscreteScheduler(num_train_timesteps, shift) as the training scripts index it
    (``noise_scheduler_copy.sigmas`` / ``.timesteps``, train_lightcontrol.py:697-702): descending, sigma = shift s / (1 + (shift-1) s)."""
    s = torch.linspace(1.0, 1.0 / num_train_timesteps, num_train_timesteps, dtype=torch.float32)
    return shift * s / (1 + (shift - 1) * s)


def flow_matching_inputs(vae, pixel_values, generator=None, num_train_timesteps: int = 1000, shift: float = 3.0, sigmas=None):
    """Steps 1 of the module docstring.  Returns (packed noisy latents [B, L, 64] bf16, timesteps [B] fp32 in [0, 1000],
    target [B, 16, h, w] fp32, latent height, latent width)."""
    with torch.no_grad():
        z = vae.encode(pixel_values).latent_dist.sample(generator=generator)
        z = ((z.float() - vae.config.shift_factor) * vae.config.scaling_factor).to(torch.bfloat16)
        B, C, h, w = z.shape
        noise = torch.randn(z.shape, generator=generator, device=z.device, dtype=torch.float32).to(torch.bfloat16)
        if sigmas is None:
            table = train_sigmas(num_train_timesteps, shift).to(z.device)
            u = compute_density_for_timestep_sampling("logit_normal", B, 0.0, 1.0, generator=generator, device=z.device)
            idx = (u * num_train_timesteps).long().clamp_(0, num_train_timesteps - 1)
            sigmas = table[idx]
        sigmas = sigmas.to(z.device, torch.float32).reshape(B)
        sg = sigmas.view(B, 1, 1, 1)
        noisy = ((1.0 - sg) * z.float() + sg * noise.float()).to(torch.bfloat16)
        packed = FluxPipeline._pack_latents(noisy, B, C, h, w)
        target = noise.float() - z.float()
    return packed, sigmas * num_train_timesteps, target, h, w


def lightcontrol_step(control_nets, transformer, vae, batch: Dict[str, torch.Tensor], optimizer=None, lr_scheduler=None,
                      max_grad_norm: float = 1.0, guidance_scale: float = 3.5, generator=None, sigmas=None, group=None):
    """One LightControl train step on this rank's batch shard.

    batch: ``pixel_values`` [B, 3, H, W] in [-1, 1] (the style image: VAE input AND control hint, train_lightcontrol.py:676,:740),
    ``prompt_embeds`` [B, S, 4096], ``pooled_prompt_embeds`` [B, 768] (outputs of the frozen MLLM + projector).
    Returns the detached loss of this rank.  control_nets: nn.ModuleList of x2i_b200 ControlNeXtModel in train mode."""
    pixel_values = batch["pixel_values"]
    if not pixel_values.is_cuda:
        raise X2IError("lightcontrol_step: CUDA tensors required (x2i_b200 has no CPU path)")
    dev = pixel_values.device
    packed, timesteps, target, h, w = flow_matching_inputs(vae, pixel_values, generator=generator, sigmas=sigmas)
    B = packed.shape[0]
    S = batch["prompt_embeds"].shape[1]
    img_ids = FluxPipeline._prepare_latent_image_ids(B, h, w, dev, torch.bfloat16)
    txt_ids = torch.zeros(S, 3, device=dev, dtype=torch.bfloat16)
    guidance = torch.full((B,), guidance_scale, device=dev, dtype=torch.float32) if transformer.config.guidance_embeds else None
    pred = transformer(hidden_states=packed, timestep=(timesteps / 1000).to(torch.bfloat16), guidance=guidance,
                       pooled_projections=batch["pooled_prompt_embeds"].to(torch.bfloat16), encoder_hidden_states=batch["prompt_embeds"].to(torch.bfloat16),
                       txt_ids=txt_ids, img_ids=img_ids, guided_hint=pixel_values.to(torch.bfloat16), control_nets=control_nets,
                       return_dict=False)[0]
    scale = 2 ** len(vae.config.block_out_channels)
    pred = FluxPipeline._unpack_latents(pred, h * scale // 2, w * scale // 2, scale)     # train_lightcontrol.py:746-751
    loss = ((pred.float() - target) ** 2).reshape(B, -1).mean(1).mean()                  # weighting "none" (:753-763)
    loss.backward()
    params = [p for p in control_nets.parameters() if p.requires_grad]
    xdist.allreduce_mean_grads_(params, group=group)                                      # the step's only collective
    if optimizer is not None:
        if max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
        optimizer.step()
        if lr_scheduler is not None:
            lr_scheduler.step()
        optimizer.zero_grad(set_to_none=True)
    return loss.detach()


def synthetic_batch(B: int, device, height: int = 1024, width: int = 1024, S: int = 512, seed: int = 0):
    """SURVEY.md 8(d) C5 shapes: style image U(-1, 1), N(0, 1) embeddings."""
    g = torch.Generator(device=device).manual_seed(seed)
    return dict(pixel_values=(torch.rand(B, 3, height, width, device=device, generator=g) * 2 - 1).bfloat16(),
                prompt_embeds=torch.randn(B, S, 4096, device=device, generator=g).bfloat16(),
                pooled_prompt_embeds=torch.randn(B, 768, device=device, generator=g).bfloat16())
