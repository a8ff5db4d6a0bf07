"""CPU/PyTorch ORACLE of the X2I hot path (FLUX MMDiT denoise step).

TEST INFRASTRUCTURE ONLY.  Nothing under ``x2i_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

What it restates
----------------
* Block / transformer structure: ``/root/reference/lightcontrol/lightcontrol_flux.py``
  (``FluxSingleTransformerBlock`` :45-104, ``FluxTransformerBlock`` :108-204,
  ``FluxTransformer2DModel`` ctor :229-284, forward :390-553).  These are in-tree
  and the oracle is PINNED to them by ``oracle/make_golden.py`` (the reference
  classes are imported there with the leaves below injected as ``diffusers``).
* Leaves (``Attention``/``FluxAttnProcessor2_0``, ``AdaLayerNorm*``, ``FeedForward``,
  ``FluxPosEmbed``, ``CombinedTimestep*Embeddings``, ``FlowMatchEulerDiscreteScheduler``)
  live in the third-party dependency ``diffusers==0.31.0``
  (``/root/reference/requirements.txt:3``) which is NOT vendored and not installable
  here.  They are restated from its published algorithm and
  cross-checked against the BFL-derived Flux in ``torchtitan`` (make_golden.py).
  PARITY OF THE LEAVES IS THEREFORE UNPINNED by the reference itself: the
  reference holds no golden vectors or tests for this path.

State-dict keys follow the diffusers FLUX checkpoint layout so
that a real ``FluxTransformer2DModel`` checkpoint loads into either this oracle
or the product module.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- leaves
class Timesteps(nn.Module):
    """Sinusoidal embedding, flip_sin_to_cos=True, downscale_freq_shift=0 (A.6)."""

    def __init__(self, num_channels: int = 256, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        expo = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device)
        expo = expo / (half - self.downscale_freq_shift)
        ang = t[:, None].float() * torch.exp(expo)[None, :]
        emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class PixArtAlphaTextProjection(nn.Module):
    def __init__(self, in_features: int, hidden_size: int, act_fn: str = "silu"):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.act_1 = nn.SiLU()
        self.linear_2 = nn.Linear(hidden_size, hidden_size)

    def forward(self, x):
        return self.linear_2(self.act_1(self.linear_1(x)))


class CombinedTimestepTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim: int, pooled_projection_dim: int):
        super().__init__()
        self.time_proj = Timesteps(256, True, 0)
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, pooled_projection):
        t = self.timestep_embedder(self.time_proj(timestep).to(pooled_projection.dtype))
        return t + self.text_embedder(pooled_projection)


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, embedding_dim: int, pooled_projection_dim: int):
        super().__init__()
        self.time_proj = Timesteps(256, True, 0)
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        self.guidance_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim)

    def forward(self, timestep, guidance, pooled_projection):
        t = self.timestep_embedder(self.time_proj(timestep).to(pooled_projection.dtype))
        g = self.guidance_embedder(self.time_proj(guidance).to(pooled_projection.dtype))
        return (t + g) + self.text_embedder(pooled_projection)


def rope_table(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """cos/sin table [L, sum(axes_dim)] fp32 from integer position ids [L, 3] (A.4).

    Frequencies are formed in float64 and the angles' cos/sin rounded to fp32,
    each value repeated for the (even, odd) pair it rotates.
    """
    pos = ids.float()
    cos_parts, sin_parts = [], []
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device)[: d // 2] / d))
        ang = torch.outer(pos[:, i].to(torch.float64), freqs)
        cos_parts.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_parts.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_parts, dim=-1), torch.cat(sin_parts, dim=-1)


class FluxPosEmbed(nn.Module):
    def __init__(self, theta: int = 10000, axes_dim=(16, 56, 56)):
        super().__init__()
        self.theta = theta
        self.axes_dim = tuple(axes_dim)

    def forward(self, ids: torch.Tensor):
        return rope_table(ids, self.axes_dim, float(self.theta))


def apply_rotary_emb(x: torch.Tensor, freqs: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """Adjacent-pair rotation of x[B,H,L,D] in fp32, result cast back (A.4)."""
    cos, sin = freqs
    cos = cos[None, None].to(x.device)
    sin = sin[None, None].to(x.device)
    pairs = x.reshape(*x.shape[:-1], -1, 2)
    even, odd = pairs.unbind(-1)
    rot = torch.stack([-odd, even], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        var = x.float().pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)  # promotes to fp32
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        return x * self.weight


class AdaLayerNormZero(nn.Module):
    def __init__(self, embedding_dim: int):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, 6 * embedding_dim)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    def __init__(self, embedding_dim: int):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, 3 * embedding_dim)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(self.silu(emb))
        shift, scale, gate = emb.chunk(3, dim=1)
        x = self.norm(x) * (1 + scale[:, None]) + shift[:, None]
        return x, gate


class AdaLayerNormContinuous(nn.Module):
    def __init__(self, embedding_dim, conditioning_embedding_dim, elementwise_affine=False, eps=1e-6):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(conditioning_embedding_dim, 2 * embedding_dim)
        self.norm = nn.LayerNorm(embedding_dim, eps=eps, elementwise_affine=elementwise_affine)

    def forward(self, x, conditioning_embedding):
        emb = self.linear(self.silu(conditioning_embedding).to(x.dtype))
        scale, shift = emb.chunk(2, dim=1)  # scale FIRST (A.1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


class GELU(nn.Module):
    """Linear followed by GELU (diffusers' activations.GELU)."""

    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, activation_fn: str = "gelu-approximate"):
        super().__init__()
        assert activation_fn == "gelu-approximate"
        self.net = nn.ModuleList([GELU(dim, dim * mult, "tanh"), nn.Dropout(0.0), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class FluxAttnProcessor2_0:
    """Joint (txt-first) attention with per-head q/k RMSNorm and RoPE (A.3)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, image_rotary_emb=None):
        B = hidden_states.shape[0]
        H = attn.heads

        def heads(t):
            return t.view(B, -1, H, t.shape[-1] // H).transpose(1, 2)

        q = attn.norm_q(heads(attn.to_q(hidden_states)))
        k = attn.norm_k(heads(attn.to_k(hidden_states)))
        v = heads(attn.to_v(hidden_states))
        if encoder_hidden_states is not None:
            eq = attn.norm_added_q(heads(attn.add_q_proj(encoder_hidden_states)))
            ek = attn.norm_added_k(heads(attn.add_k_proj(encoder_hidden_states)))
            ev = heads(attn.add_v_proj(encoder_hidden_states))
            q = torch.cat([eq, q], dim=2)
            k = torch.cat([ek, k], dim=2)
            v = torch.cat([ev, v], dim=2)
        if image_rotary_emb is not None:
            q = apply_rotary_emb(q, image_rotary_emb)
            k = apply_rotary_emb(k, image_rotary_emb)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, H * o.shape[-1]).to(q.dtype)
        if encoder_hidden_states is not None:
            s_txt = encoder_hidden_states.shape[1]
            txt, img = o[:, :s_txt], o[:, s_txt:]
            img = attn.to_out[1](attn.to_out[0](img))
            txt = attn.to_add_out(txt)
            return img, txt
        return o


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, added_kv_proj_dim=None, dim_head=128, heads=24,
                 out_dim=None, context_pre_only=None, bias=True, processor=None, qk_norm="rms_norm", eps=1e-6,
                 pre_only=False):
        super().__init__()
        inner = out_dim if out_dim is not None else dim_head * heads
        self.heads = inner // dim_head
        self.inner_dim = inner
        self.pre_only = pre_only
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.norm_q = RMSNorm(dim_head, eps)
        self.norm_k = RMSNorm(dim_head, eps)
        if added_kv_proj_dim is not None:
            self.add_q_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            self.add_k_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, inner, bias=True)
            self.norm_added_q = RMSNorm(dim_head, eps)
            self.norm_added_k = RMSNorm(dim_head, eps)
            self.to_add_out = nn.Linear(inner, query_dim, bias=True)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.processor = processor if processor is not None else FluxAttnProcessor2_0()

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


# --------------------------------------------------------------------------- blocks
class FluxSingleTransformerBlock(nn.Module):
    """lightcontrol_flux.py:45-104."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, mlp_ratio=4.0):
        super().__init__()
        self.mlp_hidden_dim = int(dim * mlp_ratio)
        self.norm = AdaLayerNormZeroSingle(dim)
        self.proj_mlp = nn.Linear(dim, self.mlp_hidden_dim)
        self.act_mlp = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(dim + self.mlp_hidden_dim, dim)
        self.attn = Attention(query_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads, out_dim=dim,
                              bias=True, qk_norm="rms_norm", eps=1e-6, pre_only=True)

    def forward(self, hidden_states, temb, image_rotary_emb=None):
        n, gate = self.norm(hidden_states, emb=temb)
        m = self.act_mlp(self.proj_mlp(n))
        a = self.attn(hidden_states=n, image_rotary_emb=image_rotary_emb)
        y = self.proj_out(torch.cat([a, m], dim=2))
        return hidden_states + gate.unsqueeze(1) * y


class FluxTransformerBlock(nn.Module):
    """lightcontrol_flux.py:108-204."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, qk_norm="rms_norm", eps=1e-6):
        super().__init__()
        self.norm1 = AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormZero(dim)
        self.attn = Attention(query_dim=dim, added_kv_proj_dim=dim, dim_head=attention_head_dim,
                              heads=num_attention_heads, out_dim=dim, context_pre_only=False, bias=True,
                              qk_norm=qk_norm, eps=eps)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim=dim, dim_out=dim)
        self.norm2_context = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff_context = FeedForward(dim=dim, dim_out=dim)

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None):
        nx, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.norm1(hidden_states, emb=temb)
        nc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = self.norm1_context(encoder_hidden_states, emb=temb)
        a_img, a_txt = self.attn(hidden_states=nx, encoder_hidden_states=nc, image_rotary_emb=image_rotary_emb)

        x = hidden_states + gate_msa.unsqueeze(1) * a_img
        nx = self.norm2(x) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        x = x + gate_mlp.unsqueeze(1) * self.ff(nx)

        c = encoder_hidden_states + c_gate_msa.unsqueeze(1) * a_txt
        nc = self.norm2_context(c) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        c = c + c_gate_mlp.unsqueeze(1) * self.ff_context(nc)
        return c, x


class FluxTransformer2DModel(nn.Module):
    """lightcontrol_flux.py:208-553; the ControlNeXt injection (:504-507) runs when control_nets is given."""

    def __init__(self, patch_size=1, in_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128,
                 num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=False,
                 axes_dims_rope=(16, 56, 56)):
        super().__init__()
        self.config = SimpleNamespace(patch_size=patch_size, in_channels=in_channels, num_layers=num_layers,
                                      num_single_layers=num_single_layers, attention_head_dim=attention_head_dim,
                                      num_attention_heads=num_attention_heads, joint_attention_dim=joint_attention_dim,
                                      pooled_projection_dim=pooled_projection_dim, guidance_embeds=guidance_embeds,
                                      axes_dims_rope=tuple(axes_dims_rope))
        self.out_channels = in_channels
        self.inner_dim = num_attention_heads * attention_head_dim
        self.pos_embed = FluxPosEmbed(theta=10000, axes_dim=axes_dims_rope)
        cls = CombinedTimestepGuidanceTextProjEmbeddings if guidance_embeds else CombinedTimestepTextProjEmbeddings
        self.time_text_embed = cls(embedding_dim=self.inner_dim, pooled_projection_dim=pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, self.inner_dim)
        self.x_embedder = nn.Linear(in_channels, self.inner_dim)
        self.transformer_blocks = nn.ModuleList(
            [FluxTransformerBlock(self.inner_dim, num_attention_heads, attention_head_dim) for _ in range(num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(self.inner_dim, num_attention_heads, attention_head_dim)
             for _ in range(num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(self.inner_dim, self.inner_dim, elementwise_affine=False, eps=1e-6)
        self.proj_out = nn.Linear(self.inner_dim, patch_size * patch_size * self.out_channels, bias=True)

    @property
    def dtype(self):
        return self.x_embedder.weight.dtype

    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None,
                img_ids=None, txt_ids=None, guidance=None, joint_attention_kwargs=None, return_dict=True,
                guided_hint=None, control_nets=None):
        x = self.x_embedder(hidden_states)
        timestep = timestep.to(x.dtype) * 1000
        if guidance is not None:
            guidance = guidance.to(x.dtype) * 1000
            temb = self.time_text_embed(timestep, guidance, pooled_projections)
        else:
            temb = self.time_text_embed(timestep, pooled_projections)
        c = self.context_embedder(encoder_hidden_states)
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        rope = self.pos_embed(torch.cat((txt_ids, img_ids), dim=0))
        for i, blk in enumerate(self.transformer_blocks):
            c, x = blk(hidden_states=x, encoder_hidden_states=c, temb=temb, image_rotary_emb=rope)
            if control_nets is not None and i < len(control_nets):  # LightControl injection, lightcontrol_flux.py:504-507
                control = control_nets[i](guided_hint, timestep)    # note: timestep is already x1000 here (:447)
                x = x + control["out"].flatten(2).transpose(1, 2).to(x.dtype) * control["scale"]
        h = torch.cat([c, x], dim=1)
        for blk in self.single_transformer_blocks:
            h = blk(hidden_states=h, temb=temb, image_rotary_emb=rope)
        h = h[:, c.shape[1]:, ...]
        out = self.proj_out(self.norm_out(h, temb))
        if not return_dict:
            return (out,)  # diffusers returns a 1-tuple (callers index [0], train_qwenvl.py:587)
        return SimpleNamespace(sample=out)


def init_synthetic_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """Deterministic synthetic weights: W~N(0,std^2), small random biases,
    RMSNorm weights near 1.  Generated on CPU in fp32 per-parameter so that the same
    values are obtained regardless of the module's device/dtype."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim >= 2:
                v = torch.randn(p.shape, generator=g) * std
            elif "norm" in name and name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = torch.randn(p.shape, generator=g) * std
            p.copy_(v.to(p.dtype))
    return model


# --------------------------------------------------------------------------- pipeline helpers (A.7, a9, a10)
def prepare_latent_image_ids(height: int, width: int) -> torch.Tensor:
    """ids[(r*W+c)] = (0, r, c) over the packed (height//2, width//2) grid (train_qwenvl.py:216-227)."""
    h2, w2 = height // 2, width // 2
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = torch.arange(h2)[:, None].float()
    ids[..., 2] = torch.arange(w2)[None, :].float()
    return ids.reshape(h2 * w2, 3)


def pack_latents(latents: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,(H/2)(W/2),4C] (train_qwenvl.py:229-234)."""
    B, C, H, W = latents.shape
    x = latents.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, (H // 2) * (W // 2), C * 4)


def unpack_latents(latents: torch.Tensor, height: int, width: int, vae_scale_factor: int) -> torch.Tensor:
    """Inverse of pack_latents (lightcontrol/train_lightcontrol.py:403-410)."""
    B, _, ch = latents.shape
    h = height // vae_scale_factor
    w = width // vae_scale_factor
    x = latents.view(B, h, w, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(B, ch // 4, h * 2, w * 2)


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.16):
    """train_qwenvl.py:236-246."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


def flow_match_sigmas(num_steps: int, mu: Optional[float] = None, shift: float = 1.0,
                      use_dynamic_shifting: bool = False) -> torch.Tensor:
    """FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(1,1/N,N), mu=) (A.7); returns N+1 sigmas."""
    import numpy as np

    s = np.linspace(1.0, 1.0 / num_steps, num_steps)
    if use_dynamic_shifting:
        s = math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0))
    else:
        s = shift * s / (1.0 + (shift - 1.0) * s)
    s = torch.from_numpy(s).to(torch.float32)
    return torch.cat([s, torch.zeros(1)])


def euler_step(x: torch.Tensor, v: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """x <- (x.float() + (sigma_next - sigma) * v).to(v.dtype) (A.7)."""
    return (x.float() + (sigma_next - sigma) * v).to(v.dtype)


@torch.no_grad()
def denoise(model: FluxTransformer2DModel, latents, prompt_embeds, pooled, height_lat, width_lat, num_steps,
            guidance_scale: float = 3.5, dynamic_shift: bool = False, emulate_bf16_time: bool = False):
    """The FluxPipeline hot loop with vae=None, output_type='latent' (A.7).

    emulate_bf16_time: when this fp32 oracle stands in for the reference's bf16 run, feed it the timestep / guidance
    values the bf16 run effectively uses (t.to(bf16) / 1000 in bf16, then * 1000 in bf16 inside the transformer,
    lightcontrol_flux.py:447-449) -- e.g. sigma 0.75 -> 752 instead of 750."""
    B, L_img, _ = latents.shape
    img_ids = prepare_latent_image_ids(height_lat, width_lat).to(latents.device, latents.dtype)
    txt_ids = torch.zeros(prompt_embeds.shape[1], 3, device=latents.device, dtype=latents.dtype)
    mu = calculate_shift(L_img, 256, 4096, 0.5, 1.16) if dynamic_shift else None
    sig = flow_match_sigmas(num_steps, mu, 1.0, dynamic_shift)
    g = None
    if model.config.guidance_embeds:
        g = torch.full((B,), guidance_scale, device=latents.device, dtype=torch.float32)
        if emulate_bf16_time:
            g = (g.to(torch.bfloat16) * 1000).float() / 1000
    for i in range(num_steps):
        t = (sig[i] * 1000).expand(B).to(latents.device, latents.dtype)
        if emulate_bf16_time:
            t = (((sig[i] * 1000).expand(B).to(torch.bfloat16) / 1000) * 1000).float().to(latents.device)
        v = model(hidden_states=latents, timestep=t / 1000, guidance=g, pooled_projections=pooled,
                  encoder_hidden_states=prompt_embeds, txt_ids=txt_ids, img_ids=img_ids, return_dict=False)[0]
        latents = euler_step(latents, v, float(sig[i]), float(sig[i + 1]))
    return latents
