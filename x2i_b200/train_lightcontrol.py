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


def train_sigmas(num_train_timesteps: int = 1000, shift: float = 3.0, use_dynamic_shifting: bool = True) -> torch.Tensor:
    """The sigma table the training script indexes (``noise_scheduler_copy.sigmas`` / ``.timesteps``, train_lightcontrol.py:697-702),
    descending.  The reference builds the scheduler with ``from_pretrained(FLUX.1-dev, subfolder="scheduler")`` (:495-499); that
    config [D031, recalled] has ``use_dynamic_shifting = true`` (shift 3.0), and diffusers' ``__init__`` applies the static shift
    ``shift s / (1 + (shift - 1) s)`` ONLY when dynamic shifting is off -- so the reference trains on the UNSHIFTED linear table
    ``s = 1, 0.999, ..., 0.001``, which is the default here.  Pass ``use_dynamic_shifting=False`` for a schnell-style scheduler."""
    s = torch.linspace(1.0, 1.0 / num_train_timesteps, num_train_timesteps, dtype=torch.float32)
    if use_dynamic_shifting:
        return s
    return shift * s / (1 + (shift - 1) * s)


def train_sigmas_from_config(config) -> torch.Tensor:
    """``train_sigmas`` for a ``scheduler_config.json`` dict / namespace (what ``FlowMatchEulerDiscreteScheduler.from_pretrained`` reads)."""
    cfg = dict(config) if isinstance(config, dict) else vars(config)
    return train_sigmas(int(cfg.get("num_train_timesteps", 1000)), float(cfg.get("shift", 1.0)), bool(cfg.get("use_dynamic_shifting", False)))


def flow_matching_inputs(vae, pixel_values, generator=None, num_train_timesteps: int = 1000, shift: float = 3.0, sigmas=None,
                         use_dynamic_shifting: bool = True):
    """Steps 1 of the module docstring.  Returns (packed noisy latents [B, L, 64] bf16, timesteps [B] fp32 in [0, 1000],
    target [B, 16, h, w] fp32, latent height, latent width)."""
    with torch.no_grad():
        z = vae.encode(pixel_values).latent_dist.sample(generator=generator)
        z = ((z.float() - vae.config.shift_factor) * vae.config.scaling_factor).to(torch.bfloat16)
        B, C, h, w = z.shape
        noise = torch.randn(z.shape, generator=generator, device=z.device, dtype=torch.float32).to(torch.bfloat16)
        if sigmas is None:
            table = train_sigmas(num_train_timesteps, shift, use_dynamic_shifting).to(z.device)
            u = compute_density_for_timestep_sampling("logit_normal", B, 0.0, 1.0, generator=generator, device=z.device)
            idx = (u * num_train_timesteps).long().clamp_(0, num_train_timesteps - 1)
            sigmas = table[idx]
        sigmas = sigmas.to(z.device, torch.float32).reshape(B)
        sg = sigmas.view(B, 1, 1, 1)
        noisy = ((1.0 - sg) * z.float() + sg * noise.float()).to(torch.bfloat16)
        packed = FluxPipeline._pack_latents(noisy, B, C, h, w)
        target = noise.float() - z.float()
    return packed, sigmas * num_train_timesteps, target, h, w


class MasterWeightOptimizer:
    """fp32 master weights + fp32 optimizer state for bf16 model parameters -- what the reference's DeepSpeed ZeRO-2 bf16 engine
    keeps (lightcontrol/accelerate_config_debug.yaml: ``zero_stage: 2``, ``mixed_precision: bf16``).  The x2i_b200 control nets hold
    bf16 parameters (the kernels read them directly); an Adam update at the reference's lr = 1e-5 is below half a bf16 ulp of a
    typical 0.02-0.05 weight, so stepping the optimizer on the bf16 tensors themselves would round most updates away.  Here the
    wrapped optimizer owns fp32 copies: ``step()`` feeds it the bf16 gradients in fp32, updates the masters and writes them back
    rounded to bf16.  Duck-types ``torch.optim.Optimizer`` for ``lightcontrol_step`` and LR schedulers (``.optimizer``)."""

    def __init__(self, params, optimizer_cls=None, **kwargs):
        self.model_params = [p for p in params if p.requires_grad]
        self.master = [p.detach().float().clone().requires_grad_(True) for p in self.model_params]
        cls = optimizer_cls if optimizer_cls is not None else torch.optim.AdamW
        self.optimizer = cls(self.master, **kwargs)
        self.param_groups = self.optimizer.param_groups

    @torch.no_grad()
    def step(self):
        for m, p in zip(self.master, self.model_params):
            m.grad = None if p.grad is None else p.grad.float()
        self.optimizer.step()
        torch._foreach_copy_(self.model_params, self.master)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model_params:
            p.grad = None
        for m in self.master:
            m.grad = None

    def state_dict(self):
        return {"optimizer": self.optimizer.state_dict(), "master": [m.detach().cpu() for m in self.master]}

    def load_state_dict(self, sd):
        self.optimizer.load_state_dict(sd["optimizer"])
        with torch.no_grad():
            for m, v in zip(self.master, sd["master"]):
                m.copy_(v)
            torch._foreach_copy_(self.model_params, self.master)


def lightcontrol_step(control_nets, transformer, vae, batch: Dict[str, torch.Tensor], optimizer=None, lr_scheduler=None,
                      max_grad_norm: float = 1.0, guidance_scale: float = 3.5, generator=None, sigmas=None, group=None):
    """One LightControl train step on this rank's batch shard.

    batch: ``pixel_values`` [B, 3, H, W] in [-1, 1] (the style image: VAE input AND control hint, train_lightcontrol.py:676,:740),
    ``prompt_embeds`` [B, S, 4096], ``pooled_prompt_embeds`` [B, 768] (outputs of the frozen MLLM + projector).
    Returns the detached loss of this rank.  control_nets: nn.ModuleList of x2i_b200 ControlNeXtModel in train mode.
    optimizer: use ``MasterWeightOptimizer(control_nets.parameters(), lr=1e-5)`` (fp32 masters, like the reference's DeepSpeed bf16
    engine); a plain optimizer on the bf16 parameters loses most updates at the reference's learning rate."""
    pixel_values = batch["pixel_values"]
    if not pixel_values.is_cuda:
        raise X2IError("lightcontrol_step: CUDA tensors required (x2i_b200 has no CPU path)")
    dev = pixel_values.device
    packed, timesteps, target, h, w = flow_matching_inputs(vae, pixel_values, generator=generator, sigmas=sigmas)
    B = packed.shape[0]
    S = batch["prompt_embeds"].shape[1]
    from .train import _cached_ids
    txt_ids, img_ids = _cached_ids(S, h, w, dev, torch.bfloat16)  # stable tensors: the RoPE cache keys on their addresses
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
