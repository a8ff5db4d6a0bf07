"""Drop-in ``FluxTransformer2DModel`` (the X2I MMDiT denoiser) on hand-written sm_100a kernels.

Mirrors the interface the reference consumes:
  * class / attribute / state-dict names of diffusers' FLUX transformer, as used by
    ``/root/reference/lightcontrol/lightcontrol_flux.py:208-553`` and loaded by
    ``train/train_qwenvl.py:417-429`` / ``infer/inference_qwenvl.py:72-75``;
  * ``forward(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
    joint_attention_kwargs, return_dict)`` -> 1-tuple / object with ``.sample``;
  * ``blk.attn`` is a real ``nn.Module`` invoked through ``__call__`` so ``register_forward_hook`` works and sees
    ``(img_out, txt_out)`` for double blocks and the raw attention output for single blocks
    (``train/train_qwenvl.py:186-214``);
  * the attention-processor plugin protocol: ``attn_processors`` / ``set_attn_processor`` /
    ``Attention.set_processor`` with ``proc(attn, hidden_states, encoder_hidden_states, attention_mask,
    image_rotary_emb)`` (``lightcontrol_flux.py:286-384``).

Every FLOP of the forward runs in libx2i_b200.so (tcgen05 GEMMs with fused epilogues, the fused attention kernel and
the row-wise kernels); PyTorch only owns the buffers.  When an input requires grad (the student pass of distillation
training) the forward runs in a saving mode and the backward through all blocks runs on the hand-written backward
kernels (flux_train.py).  There is no CPU or eager fallback.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import os

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import X2IError

BF16 = torch.bfloat16


def _no_grad_needed(*tensors):
    """Stand-alone leaves / blocks and plug-in processors are forward-only: autograd runs through the WHOLE transformer as one
    node (flux_train.FluxTrainFn, used automatically by FluxTransformer2DModel.forward when an input requires grad)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise X2IError("x2i_b200.flux: this entry point is forward-only; gradients flow through FluxTransformer2DModel.forward "
                       "(one autograd node for the whole transformer), not through a stand-alone block or a plug-in processor")


# ------------------------------------------------------------------------------------------------ leaves
class _Holder(nn.Module):
    """Parameter holders are never called; the owning block launches the fused kernels."""

    def forward(self, *a, **k):  # pragma: no cover
        raise X2IError(f"{type(self).__name__} is a parameter holder inside a fused x2i_b200 block")


class RMSNorm(_Holder):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class AdaLayerNormZero(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 6 * dim)


class AdaLayerNormZeroSingle(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 3 * dim)


class AdaLayerNormContinuous(_Holder):
    def __init__(self, dim, cond_dim, elementwise_affine=False, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.linear = nn.Linear(cond_dim, 2 * dim)


class GELU(_Holder):
    def __init__(self, dim_in, dim_out, approximate="tanh"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate


class FeedForward(nn.Module):
    """Linear -> GELU(tanh) -> Linear; callable stand-alone (un-gated) through the same kernels."""

    def __init__(self, dim, dim_out=None, mult=4, activation_fn="gelu-approximate"):
        super().__init__()
        if activation_fn != "gelu-approximate":
            raise X2IError("FeedForward: only activation_fn='gelu-approximate' (FLUX) is implemented")
        self.net = nn.ModuleList([GELU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        _no_grad_needed(x)
        h = ops.linear(x, self.net[0].proj.weight, self.net[0].proj.bias, act=1)
        return ops.linear(h, self.net[2].weight, self.net[2].bias)


class TimestepEmbedding(_Holder):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class PixArtAlphaTextProjection(_Holder):
    def __init__(self, in_features, hidden_size):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size)
        self.linear_2 = nn.Linear(hidden_size, hidden_size)


class CombinedTimestepTextProjEmbeddings(nn.Module):
    """temb = MLP(sinusoid(t)) [+ MLP(sinusoid(guidance))] + MLP(pooled)   (SURVEY.md A.6)."""

    has_guidance = False

    def __init__(self, embedding_dim, pooled_projection_dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, embedding_dim)
        if self.has_guidance:
            self.guidance_embedder = TimestepEmbedding(256, embedding_dim)
        self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim)

    @staticmethod
    def _mlp(mod, x, out=None, accumulate=False):
        h = ops.skinny_linear(x, mod.linear_1.weight, mod.linear_1.bias)
        return ops.skinny_linear(h, mod.linear_2.weight, mod.linear_2.bias, act_in=1, out=out, accumulate=accumulate)

    def forward(self, timestep, *rest):
        if self.has_guidance:
            guidance, pooled = rest
        else:
            (pooled,) = rest
            guidance = None
        temb = self._mlp(self.timestep_embedder, ops.timestep_sinusoid(timestep.float().contiguous()))
        if guidance is not None:
            self._mlp(self.guidance_embedder, ops.timestep_sinusoid(guidance.float().contiguous()), out=temb, accumulate=True)
        self._mlp(self.text_embedder, pooled.to(BF16).contiguous(), out=temb, accumulate=True)
        return temb


class CombinedTimestepGuidanceTextProjEmbeddings(CombinedTimestepTextProjEmbeddings):
    has_guidance = True


class FluxPosEmbed(nn.Module):
    def __init__(self, theta=10000, axes_dim=(16, 56, 56)):
        super().__init__()
        self.theta = theta
        self.axes_dim = tuple(axes_dim)

    def forward(self, ids):
        cos, sin, _ = ops.rope_table(ids.float(), self.axes_dim, float(self.theta))
        return cos, sin


def compact_rope(image_rotary_emb):
    """(cos[L,128], sin[L,128]) pair-repeated tables -> the [L,64,2] table the QKV epilogue reads."""
    cos, sin = image_rotary_emb
    return torch.stack([cos[:, ::2], sin[:, ::2]], dim=-1).float().contiguous()


# ------------------------------------------------------------------------------------------------ attention
class FusedCtx:
    """Private side-channel from a block to the default processor: lets the out-projection GEMM apply the AdaLN
    gate and the residual in its epilogue while the module still returns the un-gated tensors to forward hooks."""

    __slots__ = ("rope", "gate_img", "res_img", "gate_txt", "res_txt", "want_aux", "cat_buf", "ws")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class FluxAttnProcessor2_0:
    """Default processor: fused QKV GEMM (bias + per-head RMSNorm + RoPE epilogue, head-major, txt rows first),
    the fused tcgen05 attention kernel, and the out-projections."""

    def __call__(self, attn: "Attention", hidden_states, encoder_hidden_states=None, attention_mask=None,
                 image_rotary_emb=None, x2i_fused: Optional[FusedCtx] = None):
        if attention_mask is not None:
            raise X2IError("FluxAttnProcessor2_0: attention_mask is not supported (FLUX never passes one)")
        _no_grad_needed(hidden_states, encoder_hidden_states)
        B, L_img, D = hidden_states.shape
        H = attn.heads
        S = 0 if encoder_hidden_states is None else encoder_hidden_states.shape[1]
        L = S + L_img
        dev = hidden_states.device
        ctx = x2i_fused
        if ctx is not None:
            rope = ctx.rope
        else:
            rope = compact_rope(image_rotary_emb) if image_rotary_emb is not None else None
        ws = ctx.ws if ctx is not None else {}
        q = ws.get("q"); k = ws.get("k"); v = ws.get("v")
        if q is None or q.shape != (B, H, L, 128):
            q = torch.empty(B, H, L, 128, device=dev, dtype=BF16)
            k = torch.empty_like(q); v = torch.empty_like(q)
        attn._pack()
        x2 = hidden_states.reshape(B * L_img, D)
        if encoder_hidden_states is None:
            # single block: W = [q;k;v;(proj_mlp)], attention output straight into the [.., D+F] concat buffer
            cat_buf = ctx.cat_buf if ctx is not None else None
            w, b = (attn._w_qkv_mlp, attn._b_qkv_mlp) if cat_buf is not None else (attn._w_qkv, attn._b_qkv)
            ops.qkv_rope(x2, w, b, attn.norm_q.weight, attn.norm_k.weight, rope, q, k, v, H, L_img, 0, attn.norm_q.eps,
                         mlp=None if cat_buf is None else cat_buf[:, D:])
            if cat_buf is not None:
                out = cat_buf.view(B, L, -1)[:, :, :D]
                ops.attention(q, k, v, split=0, out1=out)
                if len(attn._forward_hooks) > 0:
                    out = out.clone()  # the concat buffer is reused by the next block; hooks keep their own copy
            else:
                _, out = ops.attention(q, k, v, split=0)
            return out
        c2 = encoder_hidden_states.reshape(B * S, D)
        ops.gemm_grouped(  # image + text QKV projections in one launch
            ops.desc_qkv_rope(x2, attn._w_qkv, attn._b_qkv, attn.norm_q.weight, attn.norm_k.weight, rope, q, k, v, H, L_img, S,
                              attn.norm_q.eps),
            ops.desc_qkv_rope(c2, attn._w_add_qkv, attn._b_add_qkv, attn.norm_added_q.weight, attn.norm_added_k.weight, rope,
                              q, k, v, H, S, 0, attn.norm_added_q.eps))
        a_txt = ws.get("a_txt"); a_img = ws.get("a_img")
        if a_txt is None or a_txt.shape != (B, S, D) or a_img.shape != (B, L_img, D):
            a_txt = torch.empty(B, S, D, device=dev, dtype=BF16)
            a_img = torch.empty(B, L_img, D, device=dev, dtype=BF16)
        ops.attention(q, k, v, split=S, out0=a_txt, out1=a_img)
        wo, wa = attn.to_out[0], attn.to_add_out
        if ctx is None or ctx.gate_img is None:
            img = ops.linear(a_img, wo.weight, wo.bias)
            txt = ops.linear(a_txt, wa.weight, wa.bias)
            return img, txt
        aux_img = torch.empty(B, L_img, D, device=dev, dtype=BF16) if ctx.want_aux else None
        aux_txt = torch.empty(B, S, D, device=dev, dtype=BF16) if ctx.want_aux else None
        ops.gemm_grouped(  # both out-projections with gate + residual epilogues in one launch
            ops.desc_gate_residual(a_img.view(B * L_img, D), wo.weight, wo.bias, ctx.gate_img, ctx.res_img, L_img, aux=aux_img),
            ops.desc_gate_residual(a_txt.view(B * S, D), wa.weight, wa.bias, ctx.gate_txt, ctx.res_txt, S, aux=aux_txt))
        return aux_img, aux_txt


FusedFluxAttnProcessor2_0 = FluxAttnProcessor2_0
AttentionProcessor = FluxAttnProcessor2_0


class Attention(nn.Module):
    """Parameter layout of diffusers' ``Attention`` as configured by FLUX."""

    def __init__(self, query_dim, cross_attention_dim=None, added_kv_proj_dim=None, dim_head=128, heads=24, out_dim=None,
                 context_pre_only=None, bias=True, processor=None, qk_norm="rms_norm", eps=1e-6, pre_only=False):
        super().__init__()
        if dim_head != 128:
            raise X2IError("Attention: the sm_100a attention kernel is specialised for head_dim 128 (FLUX)")
        inner = out_dim if out_dim is not None else dim_head * heads
        self.heads = inner // dim_head
        self.inner_dim = inner
        self.pre_only = pre_only
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.norm_q = RMSNorm(dim_head, eps)
        self.norm_k = RMSNorm(dim_head, eps)
        self.has_added = added_kv_proj_dim is not None
        if self.has_added:
            self.add_q_proj = nn.Linear(added_kv_proj_dim, inner)
            self.add_k_proj = nn.Linear(added_kv_proj_dim, inner)
            self.add_v_proj = nn.Linear(added_kv_proj_dim, inner)
            self.norm_added_q = RMSNorm(dim_head, eps)
            self.norm_added_k = RMSNorm(dim_head, eps)
            self.to_add_out = nn.Linear(inner, query_dim)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = processor if processor is not None else FluxAttnProcessor2_0()
        self._extra_w = None  # (weight, bias) of the owning single block's proj_mlp, fused behind q;k;v
        self._w_qkv = None

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self):
        return self.processor

    # -- fused weight storage: q;k;v (;proj_mlp) rows concatenated once, the original Parameters re-pointed at
    #    views of it (no extra memory, state_dict keys unchanged).  Re-done automatically after .to()/.cuda().
    @staticmethod
    def _fuse(linears):
        w = torch.cat([l.weight.data for l in linears], 0)
        b = torch.cat([l.bias.data for l in linears], 0)
        o = 0
        for l in linears:
            n = l.weight.shape[0]
            l.weight.data = w[o:o + n]
            l.bias.data = b[o:o + n]
            o += n
        return w, b

    def _pack(self):
        if self._w_qkv is not None and self._w_qkv.data_ptr() == self.to_q.weight.data_ptr():
            return
        if self.to_q.weight.dtype != BF16 or not self.to_q.weight.is_cuda:
            raise X2IError("x2i_b200 modules run in bf16 on a CUDA device: call .to('cuda', torch.bfloat16) first")
        lin = [self.to_q, self.to_k, self.to_v]
        if self._extra_w is not None:
            w, b = self._fuse(lin + [self._extra_w])
            D3 = 3 * self.inner_dim
            self._w_qkv_mlp, self._b_qkv_mlp = w, b
            self._w_qkv, self._b_qkv = w[:D3], b[:D3]
        else:
            self._w_qkv, self._b_qkv = self._fuse(lin)
        if self.has_added:
            self._w_add_qkv, self._b_add_qkv = self._fuse([self.add_q_proj, self.add_k_proj, self.add_v_proj])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


def _default_proc(attn: Attention) -> bool:
    return type(attn.processor) is FluxAttnProcessor2_0


# ------------------------------------------------------------------------------------------------ blocks
class FluxTransformerBlock(nn.Module):
    """Double-stream block (lightcontrol_flux.py:108-204)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, qk_norm="rms_norm", eps=1e-6):
        super().__init__()
        self.dim = dim
        self.norm1 = AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormZero(dim)
        self.attn = Attention(query_dim=dim, added_kv_proj_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads,
                              out_dim=dim, context_pre_only=False, bias=True, qk_norm=qk_norm, eps=eps)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim=dim, dim_out=dim)
        self.norm2_context = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff_context = FeedForward(dim=dim, dim_out=dim)

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, _mod=None, _rope=None, _ws=None):
        """hidden_states [B,L_img,D] and encoder_hidden_states [B,S,D] are updated IN PLACE and returned (c, x)."""
        _no_grad_needed(hidden_states, encoder_hidden_states, temb)
        x, c = hidden_states, encoder_hidden_states
        B, L_img, D = x.shape
        S = c.shape[1]
        ws = _ws if _ws is not None else {}
        if _mod is None:  # stand-alone use: compute this block's modulation from temb
            _mod = torch.cat([ops.skinny_linear(temb, self.norm1.linear.weight, self.norm1.linear.bias, act_in=1),
                              ops.skinny_linear(temb, self.norm1_context.linear.weight, self.norm1_context.linear.bias, act_in=1)], 1)
        mi = [_mod[:, i * D:(i + 1) * D] for i in range(6)]           # shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
        mc = [_mod[:, (6 + i) * D:(7 + i) * D] for i in range(6)]
        if _rope is None and image_rotary_emb is not None:
            _rope = compact_rope(image_rotary_emb)
        x2, c2 = x.view(B * L_img, D), c.view(B * S, D)
        nx_buf, nc_buf = ws.get("nx"), ws.get("nc")
        if nx_buf is None or nc_buf is None:
            nx_buf = torch.empty(B * L_img, D, device=x.device, dtype=BF16)
            nc_buf = torch.empty(B * S, D, device=x.device, dtype=BF16)
        ops.ln_modulate2(x2, mi[1], mi[0], L_img, nx_buf, c2, mc[1], mc[0], S, nc_buf)  # image + text stream in one launch
        nx, nc = nx_buf.view(B, L_img, D), nc_buf.view(B, S, D)
        if _default_proc(self.attn):
            ctx = FusedCtx(rope=_rope, gate_img=mi[2], res_img=x2, gate_txt=mc[2], res_txt=c2,
                           want_aux=len(self.attn._forward_hooks) > 0, ws=ws)
            self.attn(hidden_states=nx, encoder_hidden_states=nc, x2i_fused=ctx)
        else:  # plug-in processor: plain protocol, gate + residual applied afterwards
            a_img, a_txt = self.attn(hidden_states=nx, encoder_hidden_states=nc, image_rotary_emb=image_rotary_emb)
            ops.gate_residual_(x2, a_img.reshape(B * L_img, D), mi[2], L_img)
            ops.gate_residual_(c2, a_txt.reshape(B * S, D), mc[2], S)
        nx2, nc2 = ops.ln_modulate2(x2, mi[4], mi[3], L_img, nx_buf, c2, mc[4], mc[3], S, nc_buf)
        F = self.ff.net[0].proj.weight.shape[0]
        hx, hc = ws.get("ffx"), ws.get("ffc")
        if hx is None or hx.shape != (B * L_img, F) or hc.shape != (B * S, F):
            hx = torch.empty(B * L_img, F, device=x.device, dtype=BF16)
            hc = torch.empty(B * S, F, device=x.device, dtype=BF16)
        f0, f2, g0, g2 = self.ff.net[0].proj, self.ff.net[2], self.ff_context.net[0].proj, self.ff_context.net[2]
        ops.gemm_grouped(ops.desc_linear(nx2, f0.weight, f0.bias, hx, act=1), ops.desc_linear(nc2, g0.weight, g0.bias, hc, act=1))
        ops.gemm_grouped(ops.desc_gate_residual(hx, f2.weight, f2.bias, mi[5], x2, L_img),
                         ops.desc_gate_residual(hc, g2.weight, g2.bias, mc[5], c2, S))
        return c, x


class FluxSingleTransformerBlock(nn.Module):
    """Single-stream block (lightcontrol_flux.py:45-104)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, mlp_ratio=4.0):
        super().__init__()
        self.dim = dim
        self.mlp_hidden_dim = int(dim * mlp_ratio)
        self.norm = AdaLayerNormZeroSingle(dim)
        self.proj_mlp = nn.Linear(dim, self.mlp_hidden_dim)
        self.act_mlp = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(dim + self.mlp_hidden_dim, dim)
        self.attn = Attention(query_dim=dim, dim_head=attention_head_dim, heads=num_attention_heads, out_dim=dim, bias=True,
                              qk_norm="rms_norm", eps=1e-6, pre_only=True)
        object.__setattr__(self.attn, "_extra_w", self.proj_mlp)  # fused behind q;k;v (not a registered child)

    def forward(self, hidden_states, temb, image_rotary_emb=None, _mod=None, _rope=None, _ws=None):
        """hidden_states [B,L,D] is updated IN PLACE and returned."""
        _no_grad_needed(hidden_states, temb)
        h = hidden_states
        B, L, D = h.shape
        F = self.mlp_hidden_dim
        ws = _ws if _ws is not None else {}
        if _mod is None:
            _mod = ops.skinny_linear(temb, self.norm.linear.weight, self.norm.linear.bias, act_in=1)
        shift, scale, gate = (_mod[:, i * D:(i + 1) * D] for i in range(3))
        if _rope is None and image_rotary_emb is not None:
            _rope = compact_rope(image_rotary_emb)
        h2 = h.view(B * L, D)
        n = ops.ln_modulate(h2, scale, shift, L, out=ws.get("n")).view(B, L, D)
        cat_buf = ws.get("cat")
        if cat_buf is None or cat_buf.shape != (B * L, D + F):
            cat_buf = torch.empty(B * L, D + F, device=h.device, dtype=BF16)
        if _default_proc(self.attn):
            self.attn(hidden_states=n, x2i_fused=FusedCtx(rope=_rope, cat_buf=cat_buf, ws=ws))
        else:
            a = self.attn(hidden_states=n, image_rotary_emb=image_rotary_emb)
            cat_buf.view(B, L, D + F)[:, :, :D].copy_(a)
            ops.linear(n.view(B * L, D), self.proj_mlp.weight, self.proj_mlp.bias, act=1, out=cat_buf[:, D:])
        ops.linear_gate_residual(cat_buf, self.proj_out.weight, self.proj_out.bias, gate, h2, L)
        return h


# ------------------------------------------------------------------------------------------------ transformer
class FluxTransformer2DModel(nn.Module):
    """The FLUX MMDiT denoiser (lightcontrol_flux.py:208-553; diffusers 0.31.0 transformer_flux.py [D031])."""

    _supports_gradient_checkpointing = True

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
        D = self.inner_dim
        self.pos_embed = FluxPosEmbed(theta=10000, axes_dim=axes_dims_rope)
        cls = CombinedTimestepGuidanceTextProjEmbeddings if guidance_embeds else CombinedTimestepTextProjEmbeddings
        self.time_text_embed = cls(embedding_dim=D, pooled_projection_dim=pooled_projection_dim)
        self.context_embedder = nn.Linear(joint_attention_dim, D)
        self.x_embedder = nn.Linear(in_channels, D)
        self.transformer_blocks = nn.ModuleList(
            [FluxTransformerBlock(D, num_attention_heads, attention_head_dim) for _ in range(num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(D, num_attention_heads, attention_head_dim) for _ in range(num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(D, D, elementwise_affine=False, eps=1e-6)
        self.proj_out = nn.Linear(D, patch_size * patch_size * self.out_channels, bias=True)
        self.gradient_checkpointing = False
        self._w_mod = None
        self._ws: Dict = {}
        self._rope_cache = None
        self._graphs: Dict = {}

    # -- reference API surface ------------------------------------------------------------------------------
    @property
    def dtype(self):
        return self.x_embedder.weight.dtype

    @property
    def device(self):
        return self.x_embedder.weight.device

    @classmethod
    def from_config(cls, config):
        cfg = dict(config) if isinstance(config, dict) else vars(config)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @classmethod
    def synthetic(cls, config: dict, device="cuda", seed: int = 0, std: float = 0.02):
        """Random-weight model built directly on `device` in bf16 (benchmarks: no checkpoints are reachable)."""
        with torch.device("meta"):
            m = cls(**config)
        m = m.to(BF16).to_empty(device=device)
        return init_synthetic_(m, seed=seed, std=std).eval()

    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, **kw):
        """Load a diffusers FLUX transformer directory (config.json + *.safetensors / *.bin)."""
        import glob
        import json
        import os
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        model = cls(**cfg)
        sd = {}
        files = sorted(glob.glob(os.path.join(root, "*.safetensors")))
        if files:
            from safetensors.torch import load_file
            for fn in files:
                sd.update(load_file(fn))
        else:
            for fn in sorted(glob.glob(os.path.join(root, "*.bin"))):
                sd.update(torch.load(fn, map_location="cpu"))
        model.load_state_dict(sd)
        return model.to(torch_dtype) if torch_dtype is not None else model

    @property
    def attn_processors(self):
        procs = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                procs[f"{name}.processor"] = m.get_processor()
        return procs

    def set_attn_processor(self, processor):
        mods = [(n, m) for n, m in self.named_modules() if isinstance(m, Attention)]
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not "
                                 f"match the number of attention layers: {len(mods)}.")
            for n, m in mods:
                m.set_processor(processor[f"{n}.processor"])
        else:
            for _, m in mods:
                m.set_processor(processor)

    def fuse_qkv_projections(self):  # the QKV projections are always fused in this implementation
        return None

    def unfuse_qkv_projections(self):
        return None

    def enable_gradient_checkpointing(self):
        """train_lightcontrol.py:666.  The differentiable forward (flux_train.forward_save) then keeps only every block's inputs and
        re-runs the block's forward kernels inside the backward (the reference's torch.utils.checkpoint per block,
        lightcontrol_flux.py:475-494,:513-531): ~0.35 GB -> ~30 MB of saved activations per block and sample."""
        self.gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        self.gradient_checkpointing = False

    # -- packing of every AdaLN modulation linear into one [N_total, D] matrix (one GEMV launch per step) ----
    def _mod_linears(self):
        lins = []
        for b in self.transformer_blocks:
            lins += [b.norm1.linear, b.norm1_context.linear]
        lins += [b.norm.linear for b in self.single_transformer_blocks]
        lins.append(self.norm_out.linear)
        return lins

    def _pack(self):
        first = self._mod_linears()[0]
        if self._w_mod is not None and self._w_mod.data_ptr() == first.weight.data_ptr():
            return
        if self.dtype != BF16 or not first.weight.is_cuda:
            raise X2IError("FluxTransformer2DModel runs in bf16 on a CUDA device: call .to('cuda', torch.bfloat16) first "
                           "(x2i_b200 has no CPU or fp32 path)")
        self._w_mod, self._b_mod = Attention._fuse(self._mod_linears())
        for m in self.modules():
            if isinstance(m, Attention):
                m._pack()
        self._ws = {}
        self._graphs = {}

    def _workspace(self, B, S, L_img):
        key = (B, S, L_img, str(self.device))
        if self._ws.get("key") != key:
            self._graphs = {}  # captured graphs point into the old workspace
            D, H, dev = self.inner_dim, self.config.num_attention_heads, self.device
            F = self.single_transformer_blocks[0].mlp_hidden_dim if len(self.single_transformer_blocks) else 4 * D
            L = S + L_img
            e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
            self._ws = dict(key=key, q=e(B, H, L, 128), k=e(B, H, L, 128), v=e(B, H, L, 128), a_txt=e(B, S, D),
                            a_img=e(B, L_img, D), nx=e(B * L_img, D), nc=e(B * S, D), ffx=e(B * L_img, 4 * D),
                            ffc=e(B * S, 4 * D), n=e(B * L, D), cat=e(B * L, D + F), x=e(B, L_img, D), c=e(B, S, D),
                            h=e(B, L, D))
        return self._ws

    def _rope(self, txt_ids, img_ids):
        """(cos, sin) full tables and the compact table for the concatenated ids, cached.  A pointer miss with identical
        content (callers that rebuild their id tensors every call) keeps the cached table so captured graphs stay valid."""
        key = (txt_ids.shape[0], img_ids.shape[0], txt_ids.data_ptr(), img_ids.data_ptr(), txt_ids._version, img_ids._version)
        rc = self._rope_cache
        if rc is not None and rc[0] != key and rc[0][:2] == key[:2]:
            ids = torch.cat((txt_ids.float(), img_ids.float()), dim=0).to(self.device)
            if torch.equal(ids, rc[3]):  # host sync: callers on a hot path pass stable id tensors (pipeline / trainers cache theirs)
                self._rope_cache = rc = (key,) + rc[1:4] + ((txt_ids, img_ids),)
        if rc is None or rc[0] != key:
            ids = torch.cat((txt_ids.float(), img_ids.float()), dim=0).to(self.device)
            cos, sin, rope = ops.rope_table(ids, self.config.axes_dims_rope, 10000.0)
            # the key is (lengths, addresses, versions): hold the id tensors themselves so the caching allocator cannot hand their
            # addresses to different ids of the same length (e.g. 64x32 vs 32x64 latents) while this entry is alive
            self._rope_cache = rc = (key, (cos, sin), rope, ids, (txt_ids, img_ids))
        return rc[1], rc[2]

    # -- forward --------------------------------------------------------------------------------------------
    use_cuda_graph = True  # replay one captured graph per step instead of ~410 launches (inference, default processors)
    stack_control_nets = True  # LightControl: evaluate all ControlNeXt nets with one launch per layer (controlnext.ControlNeXtStack)
    _cn_stack = None

    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None,
                txt_ids=None, guidance=None, joint_attention_kwargs=None, guided_hint=None, control_nets=None, return_dict=True,
                x2i_modulation=None):
        """guided_hint / control_nets: the LightControl editing branch (lightcontrol_flux.py:400-401, :504-507): after each of the
        first len(control_nets) double blocks the image stream receives control_nets[i](guided_hint, timestep)['out']."""
        self._pack()
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        if guidance is not None and not self.config.guidance_embeds:
            guidance = None
        if img_ids.shape[0] != hidden_states.shape[1] or txt_ids.shape[0] != encoder_hidden_states.shape[1]:
            # the reference fails in apply_rotary_emb's broadcast; here the RoPE table would be read out of bounds
            raise X2IError(f"FluxTransformer2DModel: img_ids / txt_ids have {img_ids.shape[0]} / {txt_ids.shape[0]} rows but the sequences "
                           f"have {hidden_states.shape[1]} / {encoder_hidden_states.shape[1]} tokens")
        train_ctrl = (control_nets is not None and len(control_nets) > 0 and torch.is_grad_enabled()
                      and any(p.requires_grad for net in control_nets for p in net.parameters()))
        if train_ctrl:  # LightControl trainer: gradients flow through the frozen transformer into the control nets
            out = self._forward_train(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
                                      guided_hint=guided_hint, control_nets=control_nets)
        elif control_nets is not None and len(control_nets) > 0:
            _no_grad_needed(hidden_states, encoder_hidden_states, pooled_projections)
            stack = self._control_stack(control_nets) if self._graphable() else None
            if stack is not None:
                # LightControl editing step through the graph: the nets' mid features (hint-only part cached across steps) are computed
                # launch by launch, the transformer with the 19 injection convs is ONE graph replay
                with ops.nvtx("x2i.control_nets(stacked)"):
                    mids = stack.mid_features(guided_hint, timestep.to(BF16) * 1000)
                out = self._forward_graphed(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
                                            mod=x2i_modulation, mids=mids, control_nets=control_nets)
            else:
                out = self._forward_eager(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
                                          guided_hint=guided_hint, control_nets=control_nets, mod=x2i_modulation)
        elif torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                             for t in (hidden_states, encoder_hidden_states, pooled_projections)):
            out = self._forward_train(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance)
        elif self._graphable():
            out = self._forward_graphed(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                                        guidance, mod=x2i_modulation)
        else:
            out = self._forward_eager(hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                                      guidance, mod=x2i_modulation)
        if not return_dict:
            # diffusers returns a 1-tuple (callers index [0], train_qwenvl.py:587); the vendored LightControl class returns the BARE
            # tensor (lightcontrol_flux.py:549-550, consumed as a tensor at train_lightcontrol.py:745-751).  With control nets the
            # call is the LightControl form: hand back an object that serves both usages.
            return _TensorTuple((out,)) if control_nets is not None else (out,)
        return SimpleNamespace(sample=out)

    def _forward_train(self, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids, guidance, guided_hint=None,
                       control_nets=None):
        """Differentiable forward (student pass of the distillation step, train/train_qwenvl.py:578-587): one autograd node
        for the whole transformer (flux_train.FluxTrainFn); the forward hooks registered on every ``blk.attn``
        (train_qwenvl.py:206-214) are then invoked, in block order, with outputs that carry autograd history."""
        from . import flux_train
        for m in self.modules():
            if isinstance(m, Attention) and not _default_proc(m):
                raise X2IError("training through a plug-in attention processor is not supported (no backward for user code "
                               "inside the fused block); use the default FluxAttnProcessor2_0")
        controls = ()
        if control_nets is not None and len(control_nets) > 0:  # lightcontrol_flux.py:504-507 with differentiable control tokens
            t1000 = timestep.to(BF16) * 1000
            nets = list(control_nets)[:len(self.transformer_blocks)]
            if not all(hasattr(n, "forward_tokens") for n in nets):
                raise X2IError("training needs x2i_b200 ControlNeXtModel control nets (their backward runs on the x2i kernels)")
            controls = self._control_tokens_multistream(nets, guided_hint, t1000)
        outs = flux_train.FluxTrainFn.apply(self, 0, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids,
                                            guidance, *controls)
        nd, ns = len(self.transformer_blocks), len(self.single_transformer_blocks)
        out, hi, ht, hs = outs[0], outs[1:1 + nd], outs[1 + nd:1 + 2 * nd], outs[1 + 2 * nd:1 + 2 * nd + ns]
        for i, blk in enumerate(self.transformer_blocks):
            for hook in list(blk.attn._forward_hooks.values()):
                hook(blk.attn, (), (hi[i], ht[i]))
        for i, blk in enumerate(self.single_transformer_blocks):
            for hook in list(blk.attn._forward_hooks.values()):
                hook(blk.attn, (), hs[i])
        return out

    # opt-in (X2I_OVERLAP_MOD=1): the modulation GEMV of all but the first block runs on a side stream next to block 0 (see _forward_eager).
    # Measured neutral on B200 (62.17 / 62.29 ms off, 62.37 / 62.26 ms on): the 1.3 ms of HBM streaming it hides is paid back by block 0's
    # GEMMs and attention sharing SMs and L2 with it, so the default stays the single launch.
    overlap_modulation = os.environ.get("X2I_OVERLAP_MOD", "0") == "1"
    _mod_stream = None
    control_net_streams = 8  # LightControl training: the independent control nets run round-robin on this many CUDA streams (1 = off)
    _cn_streams = None

    def _control_tokens_multistream(self, nets, guided_hint, t1000):
        """Differentiable control tokens of every net.  The nets are independent of each other and of the image stream and consist of
        ~50 short kernels each (10-60 us; forward AND backward), so they are issued round-robin on a few side streams: launch ramps,
        tails and small grids of one net overlap with the others'.  Autograd runs each backward op on the stream of its forward op and
        synchronises the stream boundaries itself; op workspaces are per (op, device, stream)."""
        S = max(1, min(int(self.control_net_streams), len(nets)))
        if S == 1:
            return tuple(n.forward_tokens(guided_hint, t1000) for n in nets)
        if self._cn_streams is None or len(self._cn_streams) != S:
            self._cn_streams = [torch.cuda.Stream(device=self.device) for _ in range(S)]
        cur = torch.cuda.current_stream()
        for st in self._cn_streams:
            st.wait_stream(cur)  # hint, timestep and the nets' parameters are ready
        controls = [None] * len(nets)
        for i, n in enumerate(nets):
            with torch.cuda.stream(self._cn_streams[i % S]):
                controls[i] = n.forward_tokens(guided_hint, t1000)
        for st in self._cn_streams:
            cur.wait_stream(st)
        for c in controls:
            c.record_stream(cur)  # allocated on a side stream, consumed by the transformer on this one
        return tuple(controls)

    def _graphable(self):
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return False
        for m in self.modules():
            if isinstance(m, Attention) and (len(m._forward_hooks) > 0 or not _default_proc(m)):
                return False  # hooks / plug-in processors run Python per block: eager path
        return True

    def _control_stack(self, control_nets):
        """The ControlNeXtStack of `control_nets` (cached), or None when the nets cannot be stacked."""
        from .controlnext import ControlNeXtStack
        if not (self.stack_control_nets and ControlNeXtStack.supported(control_nets)):
            return None
        if self._cn_stack is None or self._cn_stack.nets != list(control_nets):
            self._cn_stack = ControlNeXtStack(control_nets)
        return self._cn_stack

    def _forward_graphed(self, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids, guidance, mod=None, mids=None,
                         control_nets=None):
        B, L_img, _ = hidden_states.shape
        S = encoder_hidden_states.shape[1]
        _, rope = self._rope(txt_ids, img_ids)
        key = (B, S, L_img, guidance is not None, rope.data_ptr(), self._w_mod.data_ptr(), mod is not None)
        if mids is not None:  # the injection convs read the nets' last-conv weights inside the graph
            key += tuple((n.mid_convs[1].weight.data_ptr(), n.mid_convs[1].weight._version, n.mid_convs[1].bias._version) for n in control_nets)
            key += (tuple(mids.shape),)
        st = self._graphs.get(key) if hasattr(self, "_graphs") else None
        if st is None:
            if not hasattr(self, "_graphs"):
                self._graphs = {}
            dev = self.device
            sin = dict(h=torch.empty(B, L_img, hidden_states.shape[2], device=dev, dtype=BF16),
                       e=torch.empty(B, S, encoder_hidden_states.shape[2], device=dev, dtype=BF16),
                       p=torch.empty(B, pooled.shape[1], device=dev, dtype=BF16),
                       t=torch.empty(B, device=dev, dtype=torch.float32),
                       g=torch.empty(B, device=dev, dtype=torch.float32) if guidance is not None else None,
                       m=torch.empty_like(mod) if mod is not None else None,
                       c=torch.empty_like(mids) if mids is not None else None)
            self._copy_inputs(sin, hidden_states, encoder_hidden_states, pooled, timestep, guidance, mod, mids)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up outside the capture: packing, workspaces, function attributes
                self._forward_eager(sin["h"], sin["e"], sin["p"], sin["t"], img_ids, txt_ids, sin["g"], mod=sin["m"], mids=sin["c"],
                                    control_nets=control_nets)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                out = self._forward_eager(sin["h"], sin["e"], sin["p"], sin["t"], img_ids, txt_ids, sin["g"], mod=sin["m"], mids=sin["c"],
                                          control_nets=control_nets)
            st = (graph, sin, out, _lib.launch_count() - n0)
            self._graphs = {key: st}  # keep one shape resident (24 GB of weights leave room, but workspaces are per shape)
        graph, sin, out, n_kernels = st
        self._copy_inputs(sin, hidden_states, encoder_hidden_states, pooled, timestep, guidance, mod, mids)
        graph.replay()
        _lib.note_graph_replay(n_kernels)
        return out.clone()

    @staticmethod
    def _copy_inputs(sin, hidden_states, encoder_hidden_states, pooled, timestep, guidance, mod=None, mids=None):
        sin["h"].copy_(hidden_states, non_blocking=True)
        sin["e"].copy_(encoder_hidden_states, non_blocking=True)
        sin["p"].copy_(pooled, non_blocking=True)
        sin["t"].copy_(timestep.expand(sin["t"].shape[0]) if timestep.dim() > 0 else timestep, non_blocking=True)
        if sin["g"] is not None:
            sin["g"].copy_(guidance.expand(sin["g"].shape[0]) if guidance.dim() > 0 else guidance, non_blocking=True)
        if sin.get("m") is not None:
            sin["m"].copy_(mod, non_blocking=True)
        if sin.get("c") is not None:
            sin["c"].copy_(mids, non_blocking=True)

    def precompute_modulation(self, timesteps, pooled_projections, guidance=None):
        """Every AdaLN modulation of EVERY step of a sampling schedule in one pass over the modulation weights.

        The 77 modulation linears (6.5 GB at FLUX size) depend only on (timestep, guidance, pooled text), all known before the
        first step, so a T-step call streams those weights once with T x B activation rows instead of T times with B rows
        (1.3 ms of HBM time per step at 1024 px).  Rows are independent in the skinny GEMV, so the result is bit-identical to the
        per-step computation.  timesteps: [T] in the t/1000 scale the transformer is called with; returns [T, B, n_mod] bf16,
        step i is passed back as ``forward(..., x2i_modulation=mod[i])``."""
        self._pack()
        T, B = timesteps.shape[0], pooled_projections.shape[0]
        t1000 = (timesteps.to(BF16) * 1000).repeat_interleave(B)
        pooled = pooled_projections.to(BF16).repeat(T, 1)
        out = torch.empty(T * B, self._w_mod.shape[0], device=self.device, dtype=BF16)
        for r0 in range(0, T * B, 64):  # skinny_linear takes <= 64 rows per call (8 per pass over the weights)
            r1 = min(T * B, r0 + 64)
            if guidance is not None and self.config.guidance_embeds:
                g1000 = (guidance.to(BF16) * 1000).repeat(T)[r0:r1]
                temb = self.time_text_embed(t1000[r0:r1], g1000, pooled[r0:r1])
            else:
                temb = self.time_text_embed(t1000[r0:r1], pooled[r0:r1])
            ops.skinny_linear(temb, self._w_mod, self._b_mod, act_in=1, out=out[r0:r1])
        return out.view(T, B, -1)

    def _forward_eager(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
                       guided_hint=None, control_nets=None, mod=None, mids=None):
        B, L_img, _ = hidden_states.shape
        S = encoder_hidden_states.shape[1]
        D = self.inner_dim
        ws = self._workspace(B, S, L_img)
        rope_full, rope = self._rope(txt_ids, img_ids)

        with ops.nvtx("x2i.embed+modulation"):
            x = ops.linear(hidden_states.to(BF16).contiguous(), self.x_embedder.weight, self.x_embedder.bias, out=ws["x"])
            # timestep / guidance arrive as t/1000; the reference scales them IN bf16 (lightcontrol_flux.py:447-449)
            t1000 = timestep.to(BF16) * 1000
            c = ops.linear(encoder_hidden_states.to(BF16).contiguous(), self.context_embedder.weight, self.context_embedder.bias,
                           out=ws["c"])
            temb = None
            mod_join = None
            if mod is None:  # otherwise: precompute_modulation() already produced this step's rows
                if guidance is not None:
                    temb = self.time_text_embed(t1000, guidance.to(BF16) * 1000, pooled_projections)
                else:
                    temb = self.time_text_embed(t1000, pooled_projections)
                # every AdaLN modulation of this step.  The GEMV streams 6.5 GB of weights (HBM-bound, ~1.3 ms) and only the first double
                # block needs its rows right away: its 12 * D rows are computed here, the rest on a side stream while block 0's
                # tensor-bound kernels run (joined before block 1; inside a graph capture this is a fork / join of the graph).
                n0 = 12 * D if (self.overlap_modulation and len(self.transformer_blocks) > 1) else 0
                if n0 == 0:
                    mod = ops.skinny_linear(temb, self._w_mod, self._b_mod, act_in=1)
                else:
                    mod = torch.empty(temb.shape[0], self._w_mod.shape[0], device=temb.device, dtype=BF16)
                    ops.skinny_linear(temb, self._w_mod[:n0], self._b_mod[:n0], act_in=1, out=mod[:, :n0])
                    if self._mod_stream is None:
                        self._mod_stream = torch.cuda.Stream(device=self.device)
                    cur_stream = torch.cuda.current_stream()
                    self._mod_stream.wait_stream(cur_stream)
                    with torch.cuda.stream(self._mod_stream):
                        ops.skinny_linear(temb, self._w_mod[n0:], self._b_mod[n0:], act_in=1, out=mod[:, n0:])
                    mod_join = self._mod_stream

        if mids is None and control_nets is not None and len(control_nets) > 0:
            stack = self._control_stack(control_nets)
            if stack is not None:
                with ops.nvtx("x2i.control_nets(stacked)"):
                    mids = stack.mid_features(guided_hint, t1000)  # all nets, one launch per layer (they do not depend on x)

        off = 0
        for i, blk in enumerate(self.transformer_blocks):
            if i == 1 and mod_join is not None:
                torch.cuda.current_stream().wait_stream(mod_join)  # the remaining modulation rows have landed
                mod_join = None
            with ops.nvtx("x2i.double_block"):
                c, x = blk(hidden_states=x, encoder_hidden_states=c, temb=temb, image_rotary_emb=rope_full,
                           _mod=mod[:, off:off + 12 * D], _rope=rope, _ws=ws)
            off += 12 * D
            if control_nets is not None and i < len(control_nets):  # lightcontrol_flux.py:504-507
                net = control_nets[i]
                if mids is not None:
                    net.finish_tokens(mids[i], add_to=x)
                elif hasattr(net, "forward_tokens"):   # x2i_b200 ControlNeXtModel: adds in the epilogue of its last conv
                    net.forward_tokens(guided_hint, t1000, add_to=x)
                else:                                # any other control net: the reference's protocol, then a fused axpy
                    control = net(guided_hint, t1000)
                    sig = control["out"].flatten(2).transpose(1, 2).to(BF16).contiguous()
                    ops.euler_step_(x, sig, float(control["scale"]))
        h = ws["h"]
        h[:, :S].copy_(c)
        h[:, S:].copy_(x)
        for blk in self.single_transformer_blocks:
            with ops.nvtx("x2i.single_block"):
                h = blk(hidden_states=h, temb=temb, image_rotary_emb=rope_full, _mod=mod[:, off:off + 3 * D], _rope=rope, _ws=ws)
            off += 3 * D
        # norm_out (AdaLayerNormContinuous: scale first, then shift) + proj_out over all rows; text rows dropped after
        with ops.nvtx("x2i.norm_out+proj_out"):
            n = ops.ln_modulate(h.view(B * (S + L_img), D), mod[:, off:off + D], mod[:, off + D:off + 2 * D], S + L_img,
                                out=ws["n"])
            return ops.linear(n, self.proj_out.weight, self.proj_out.bias).view(B, S + L_img, -1)[:, S:].contiguous()


class _TensorTuple(tuple):
    """1-tuple ``(sample,)`` that also forwards tensor attributes (``.shape``, ``.view``, ``.float()`` ...) to the sample, so both
    ``model(...)[0]`` (diffusers) and ``_unpack_latents(model(...), ...)`` (lightcontrol/train_lightcontrol.py:732-751) work."""

    def __getattr__(self, name):
        return getattr(tuple.__getitem__(self, 0), name)


def init_synthetic_(model: nn.Module, seed: int = 0, std: float = 0.02) -> nn.Module:
    """Synthetic weights for benchmarks (no checkpoints are reachable): W ~ N(0, std^2), small biases, RMSNorm ~ 1.
    Generated on the parameter's own device."""
    with torch.no_grad():
        for i, (name, p) in enumerate(model.named_parameters()):
            g = torch.Generator(device=p.device).manual_seed(seed * 100003 + i)
            if p.ndim >= 2:
                p.copy_((torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std).to(p.dtype))
            elif "norm" in name and name.endswith("weight"):
                p.copy_((1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)).to(p.dtype))
            else:
                p.copy_((torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std).to(p.dtype))
    return model
