"""Drop-in MiniCPM-o perceiver ``Resampler`` (``minicpm/resampler.py:83-189`` of the reference) on sm_100a kernels.

64 learned queries cross-attend to the SigLIP patch features of every image with a 2-D sincos positional table added to
the keys and a key-padding mask for ragged patch counts; wired in the reference at ``minicpm/modeling_minicpmo.py:205-212,
:350`` as ``Resampler(num_queries=64, embed_dim=3584, num_heads=28, kv_dim=1152, adaptive=True)``.  Same constructor,
parameter names (``query, kv_proj.weight, attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_q.*, ln_kv.*,
ln_post.*, proj``) and ``forward(x, tgt_sizes)`` contract.  Kernels: tcgen05 GEMMs (kv_proj, the three in-projections
stored head-major, out_proj, ``@ proj`` as an MN-major operand), affine LayerNorm, the positional gather/add, and the fused
attention kernel in its cross-attention form (Lq != Lkv, per-image key lengths).  Forward only; head_dim must be 128.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16


def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("hw,d->hwd", pos, omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=-1)


def get_2d_sincos_pos_embed(embed_dim, image_size):
    """[H, W, embed_dim] table, numpy fp32 like the reference (minicpm/resampler.py:32-80; 'w goes first')."""
    gh, gw = (image_size, image_size) if isinstance(image_size, int) else (image_size[0], image_size[1])
    grid = np.stack(np.meshgrid(np.arange(gw, dtype=np.float32), np.arange(gh, dtype=np.float32)), axis=0)
    return np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=-1)


class _MHAParams(nn.Module):
    """Parameter layout of torch / the reference's MultiheadAttention (packed in-projection)."""

    def __init__(self, embed_dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)


class Resampler(nn.Module):
    def __init__(self, num_queries, embed_dim, num_heads, kv_dim=None, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                 adaptive=False, max_size=(70, 70)):
        super().__init__()
        if embed_dim != num_heads * 128:
            raise X2IError("Resampler: the attention kernel is specialised for head_dim 128 (embed_dim == 128 * num_heads)")
        self.num_queries, self.embed_dim, self.num_heads = num_queries, embed_dim, num_heads
        self.adaptive, self.max_size = adaptive, list(max_size)
        self.query = nn.Parameter(torch.zeros(num_queries, embed_dim))
        self.kv_proj = nn.Linear(kv_dim, embed_dim, bias=False) if kv_dim is not None and kv_dim != embed_dim else nn.Identity()
        self.attn = _MHAParams(embed_dim)
        self.ln_q, self.ln_kv, self.ln_post = norm_layer(embed_dim), norm_layer(embed_dim), norm_layer(embed_dim)
        self.proj = nn.Parameter((embed_dim ** -0.5) * torch.randn(embed_dim, embed_dim))
        self._set_2d_pos_cache(self.max_size)

    def _set_2d_pos_cache(self, max_size, device="cpu"):
        pos = torch.from_numpy(get_2d_sincos_pos_embed(self.embed_dim, max_size)).float().to(device)
        self.register_buffer("pos_embed", pos, persistent=False)
        self._pos_bf16 = None

    def _adjust_pos_cache(self, max_h, max_w, device):
        if max_h > self.max_size[0] or max_w > self.max_size[1]:
            self.max_size = [max(max_h, self.max_size[0]), max(max_w, self.max_size[1])]
            self._set_2d_pos_cache(self.max_size, device)

    @torch.no_grad()
    def forward(self, x, tgt_sizes=None):
        assert x.shape[0] == tgt_sizes.shape[0]
        if not x.is_cuda or self.query.dtype != BF16:
            raise X2IError("Resampler runs in bf16 on a CUDA device (x2i_b200 has no CPU path): .to('cuda', torch.bfloat16)")
        B, L, _ = x.shape
        D, H, Q = self.embed_dim, self.num_heads, self.num_queries
        sizes_host = tgt_sizes.cpu()
        self._adjust_pos_cache(int(sizes_host[:, 0].max()), int(sizes_host[:, 1].max()), x.device)
        if int((sizes_host[:, 0] * sizes_host[:, 1]).max()) != L:
            raise X2IError("Resampler: x must be padded to max(h*w) patches, as the reference expects")
        if self._pos_bf16 is None or self._pos_bf16.device != x.device:
            self._pos_bf16 = self.pos_embed.to(x.device, BF16).contiguous()
        sizes = tgt_sizes.to(x.device, torch.int32).contiguous()
        kv_len = (sizes[:, 0] * sizes[:, 1]).contiguous()

        xb = x.to(BF16).contiguous()
        xk = ops.linear(xb, self.kv_proj.weight, None) if isinstance(self.kv_proj, nn.Linear) else xb
        xn = ops.layernorm_affine(xk, self.ln_kv.weight, self.ln_kv.bias, self.ln_kv.eps)
        qn = ops.layernorm_affine(self.query, self.ln_q.weight, self.ln_q.bias, self.ln_q.eps)
        kin = ops.add_pos2d(xn, self._pos_bf16, sizes)

        W, bias = self.attn.in_proj_weight, self.attn.in_proj_bias
        q1 = torch.empty(1, H, Q, 128, device=x.device, dtype=BF16)
        ops.qkv_rope(qn, W[:D], bias[:D], None, None, None, q1, None, None, H, Q, 0)
        kh = torch.empty(B, H, L, 128, device=x.device, dtype=BF16)
        vh = torch.empty_like(kh)
        ops.gemm_grouped(  # key and value in-projections (different inputs) in one launch, stored head-major
            ops.desc_qkv_rope(kin.view(B * L, D), W[D:2 * D], bias[D:2 * D], None, None, None, kh, None, None, H, L, 0),
            ops.desc_qkv_rope(xn.view(B * L, D), W[2 * D:], bias[2 * D:], None, None, None, vh, None, None, H, L, 0))
        o = ops.cross_attention(q1.expand(B, H, Q, 128).contiguous(), kh, vh, kv_len)  # [B, Q, D]
        o = ops.linear(o, self.attn.out_proj.weight, self.attn.out_proj.bias)
        o = ops.layernorm_affine(o, self.ln_post.weight, self.ln_post.bias, self.ln_post.eps)
        return ops.matmul_kn(o.view(B * Q, D), self.proj).view(B, Q, D)
