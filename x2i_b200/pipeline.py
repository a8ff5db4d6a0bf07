"""``FluxPipeline`` / ``FlowMatchEulerDiscreteScheduler`` drop-ins for the X2I inference entry points.

The reference builds ``FluxPipeline.from_pretrained(path, text_encoder=None, text_encoder_2=None, tokenizer=None,
tokenizer_2=None, vae=None)`` and calls it with pre-computed ``prompt_embeds`` / ``pooled_prompt_embeds`` and
``output_type="latent"`` (``infer/inference_qwenvl.py:72-73,:188-212``; minicpm ``:186-210``; internvl ``:196-220``;
multi_turn ``:149-162``), then unpacks with the static ``FluxPipeline._unpack_latents``.  This file provides exactly
that surface (diffusers 0.31.0 semantics) over x2i_b200.flux.FluxTransformer2DModel; the Euler update
runs in the x2i_euler_step kernel.  Prompt encoding, VAE decode and image post-processing are out of scope.
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops
from ._lib import X2IError
from .flux import FluxTransformer2DModel


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16):
    """mu = m * image_seq_len + b  (train/train_qwenvl.py:236-246)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """train/train_qwenvl.py:248-281."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    if timesteps is not None:
        raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support custom"
                         f" timestep schedules. Please check whether you are using the correct scheduler.")
    if sigmas is not None:
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, len(scheduler.timesteps)


class FlowMatchEulerDiscreteScheduler:
    """diffusers 0.31.0 ``FlowMatchEulerDiscreteScheduler`` [D031] restricted to what FluxPipeline uses."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096, **unused):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                      use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift, max_shift=max_shift,
                                      base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len)
        self.timesteps = None
        self.sigmas = None
        self._step_index = None

    @classmethod
    def from_config(cls, config):
        cfg = dict(config) if isinstance(config, dict) else vars(config)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @property
    def step_index(self):
        return self._step_index

    @staticmethod
    def time_shift(mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu: Optional[float] = None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have a pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            n = self.config.num_train_timesteps
            ts = np.linspace(n * 1.0, n * (1.0 / n), num_inference_steps)  # sigma_max=1, sigma_min=1/n for shift=1
            sigmas = ts / n
        sigmas = np.asarray(sigmas, dtype=np.float64)
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sig = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self._sigmas_host = torch.cat([sig, torch.zeros(1)])            # host copy: no device sync inside the loop
        self.timesteps = (sig * self.config.num_train_timesteps).to(device=device)
        self.sigmas = self._sigmas_host.to(device=device)
        self.num_inference_steps = len(sig)
        self._step_index = None

    def step(self, model_output, timestep=None, sample=None, return_dict: bool = True, **kw):
        """sample <- (sample.float() + (sigma_next - sigma) * model_output).to(model_output.dtype), in place when
        ``sample`` is a contiguous bf16 CUDA tensor (x2i_euler_step)."""
        if self._step_index is None:
            self._step_index = 0
        i = self._step_index
        dsigma = float(self._sigmas_host[i + 1] - self._sigmas_host[i])  # fp32 subtraction, as the reference
        if sample.dtype != torch.bfloat16 or not sample.is_contiguous():
            sample = sample.to(torch.bfloat16).contiguous()
        prev = ops.euler_step_(sample, model_output.contiguous(), dsigma)
        self._step_index += 1
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: draw on the generator's device, then move."""
    gdev = generator.device if generator is not None else torch.device(device or "cpu")
    if isinstance(generator, (list, tuple)):
        gdev = generator[0].device
        x = torch.cat([torch.randn((1,) + tuple(shape[1:]), generator=g, device=gdev, dtype=dtype) for g in generator])
    else:
        x = torch.randn(tuple(shape), generator=generator, device=gdev, dtype=dtype)
    return x.to(device)


class FluxPipeline:
    """Latent-space FLUX sampler: ``pipeline(prompt_embeds=, pooled_prompt_embeds=, num_inference_steps=,
    guidance_scale=, height=, width=, output_type="latent", generator=).images -> [B, (h/16)(w/16), 64]``."""

    hoist_modulation = True  # all steps' AdaLN modulations in one pass over the modulation weights (bit-identical; see __call__)

    def __init__(self, scheduler=None, vae=None, text_encoder=None, tokenizer=None, text_encoder_2=None, tokenizer_2=None,
                 transformer=None):
        self.scheduler = scheduler if scheduler is not None else FlowMatchEulerDiscreteScheduler()
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.text_encoder_2, self.tokenizer_2 = text_encoder_2, tokenizer_2
        self.transformer = transformer
        self.vae_scale_factor = 16  # diffusers 0.31.0 value when vae is None
        self.default_sample_size = 64
        self._device = transformer.device if transformer is not None else torch.device("cpu")
        self._ids_cache = {}

    @classmethod
    def from_pretrained(cls, path, text_encoder=None, text_encoder_2=None, tokenizer=None, tokenizer_2=None, vae=None,
                        transformer=None, torch_dtype=None, **kw):
        if transformer is None:
            transformer = FluxTransformer2DModel.from_pretrained(path, subfolder="transformer", torch_dtype=torch_dtype)
        sched_cfg = os.path.join(path, "scheduler", "scheduler_config.json")
        if os.path.exists(sched_cfg):
            with open(sched_cfg) as f:
                scheduler = FlowMatchEulerDiscreteScheduler.from_config(json.load(f))
        else:
            scheduler = FlowMatchEulerDiscreteScheduler()
        return cls(scheduler=scheduler, vae=vae, text_encoder=text_encoder, tokenizer=tokenizer,
                   text_encoder_2=text_encoder_2, tokenizer_2=tokenizer_2, transformer=transformer)

    def to(self, *args, **kwargs):
        self.transformer.to(*args, **kwargs)
        self._device = self.transformer.device
        return self

    @property
    def device(self):
        return self._device

    # ---- static helpers the reference calls on the class ------------------------------------------------------
    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        """ids[r*W+c] = (0, r, c) over the packed grid (train/train_qwenvl.py:216-227)."""
        ids = torch.zeros(height // 2, width // 2, 3)
        ids[..., 1] = ids[..., 1] + torch.arange(height // 2)[:, None]
        ids[..., 2] = ids[..., 2] + torch.arange(width // 2)[None, :]
        return ids.reshape((height // 2) * (width // 2), 3).to(device=device, dtype=dtype)

    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width):
        """[B,C,H,W] -> [B,(H/2)(W/2),4C] (train/train_qwenvl.py:229-234)."""
        latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
        latents = latents.permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        """[B,(h)(w),4C] -> [B,C,2h,2w] (lightcontrol/train_lightcontrol.py:403-410)."""
        batch_size, num_patches, channels = latents.shape
        height = height // vae_scale_factor
        width = width // vae_scale_factor
        latents = latents.view(batch_size, height, width, channels // 4, 2, 2)
        latents = latents.permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // (2 * 2), height * 2, width * 2)

    def _cached_text_ids(self, S, device, dtype):
        ck = ("txt", S, str(device), dtype)
        if ck not in self._ids_cache:
            self._ids_cache[ck] = torch.zeros(S, 3, device=device, dtype=dtype)
        return self._ids_cache[ck]

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        height = 2 * (int(height) // self.vae_scale_factor)
        width = 2 * (int(width) // self.vae_scale_factor)
        ck = ("img", height, width, str(device), dtype)
        if ck not in self._ids_cache:  # stable id tensors across calls keep the transformer's RoPE table and graph cached
            self._ids_cache[ck] = self._prepare_latent_image_ids(batch_size, height, width, device, dtype)
        ids = self._ids_cache[ck]
        if latents is not None:
            # the Euler update runs in place on the latents: never on the caller's tensor (diffusers' step returns a new one)
            return latents.to(device=device, dtype=dtype, copy=True), ids
        latents = randn_tensor((batch_size, num_channels_latents, height, width), generator, device, dtype)
        return self._pack_latents(latents, batch_size, num_channels_latents, height, width), ids

    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 28, timesteps=None, guidance_scale: float = 3.5, num_images_per_prompt: int = 1,
                 generator=None, latents=None, prompt_embeds=None, pooled_prompt_embeds=None, output_type: str = "pil",
                 return_dict: bool = True, joint_attention_kwargs=None, max_sequence_length: int = 512, guided_hint=None,
                 control_nets=None, **kw):
        """guided_hint [B,3,H,W] / control_nets (ModuleList of ControlNeXtModel): the LightControl editing branch; they are handed
        to the transformer every step exactly as lightcontrol/train_lightcontrol.py:732-743 does (the reference ships no
        LightControl inference script; BASELINE config 5 = 20 Euler steps of this call)."""
        if prompt is not None or prompt_embeds is None or pooled_prompt_embeds is None:
            raise X2IError("FluxPipeline (x2i_b200): text encoders are out of scope; pass prompt_embeds and "
                           "pooled_prompt_embeds as the X2I inference scripts do")
        if output_type != "latent" and self.vae is None:
            raise X2IError("FluxPipeline (x2i_b200): output_type != 'latent' needs a vae (x2i_b200.vae.AutoencoderKL); the X2I "
                           "scripts construct the pipeline with vae=None and decode themselves")
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        device = self.transformer.device
        dtype = prompt_embeds.dtype
        B = prompt_embeds.shape[0] * num_images_per_prompt
        if num_images_per_prompt != 1:
            prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, 0)
            pooled_prompt_embeds = pooled_prompt_embeds.repeat_interleave(num_images_per_prompt, 0)
        prompt_embeds = prompt_embeds.to(device)
        pooled_prompt_embeds = pooled_prompt_embeds.to(device)
        text_ids = self._cached_text_ids(prompt_embeds.shape[1], device, dtype)
        latents, latent_image_ids = self.prepare_latents(B, self.transformer.config.in_channels // 4, height, width, dtype,
                                                         device, generator, latents)
        latents = latents.contiguous()
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        sc = self.scheduler.config
        mu = calculate_shift(latents.shape[1], sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        ts, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps, sigmas, mu=mu)
        guidance = None
        if self.transformer.config.guidance_embeds:
            guidance = torch.full([1], guidance_scale, device=device, dtype=torch.float32).expand(B)
        # The AdaLN modulations depend only on (timestep, guidance, pooled text): all steps' rows come out of ONE pass over the 6.5 GB of
        # modulation weights (bit-identical to the per-step GEMV; FluxTransformer2DModel.precompute_modulation).  A foreign transformer
        # without that method is called the plain way.
        extra = [{} for _ in range(num_inference_steps)]
        if self.hoist_modulation and hasattr(self.transformer, "precompute_modulation"):
            mods = self.transformer.precompute_modulation(ts.to(latents.dtype) / 1000, pooled_prompt_embeds, guidance)
            extra = [dict(x2i_modulation=mods[i]) for i in range(num_inference_steps)]
        for i in range(num_inference_steps):
            timestep = ts[i].expand(B).to(latents.dtype)
            noise_pred = self.transformer(hidden_states=latents, timestep=timestep / 1000, guidance=guidance,
                                          pooled_projections=pooled_prompt_embeds, encoder_hidden_states=prompt_embeds,
                                          txt_ids=text_ids, img_ids=latent_image_ids,
                                          joint_attention_kwargs=joint_attention_kwargs, guided_hint=guided_hint,
                                          control_nets=control_nets, return_dict=False, **extra[i])[0]
            latents = self.scheduler.step(noise_pred, ts[i], latents, return_dict=False)[0]
        if output_type != "latent":  # diffusers' tail: unpack -> un-scale -> vae.decode -> image_processor.postprocess
            from .vae import VaeImageProcessor
            scale = 2 ** len(self.vae.config.block_out_channels)
            z = self._unpack_latents(latents, height, width, scale)
            z = (z / self.vae.config.scaling_factor) + self.vae.config.shift_factor
            latents = VaeImageProcessor(vae_scale_factor=scale).postprocess(self.vae.decode(z, return_dict=False)[0], output_type=output_type)
        if not return_dict:
            return (latents,)
        return SimpleNamespace(images=latents)
