"""ORACLE (test infrastructure only) for the MiniCPM-o perceiver resampler.

Restates ``/root/reference/minicpm/resampler.py``: ``get_2d_sincos_pos_embed`` :32-80 and ``Resampler`` :83-189
(forward :146-186).  The reference's private ``MultiheadAttention`` (:192-668) is a copy of torch's; torch's own
``nn.MultiheadAttention`` is used here (same parameters ``in_proj_weight / in_proj_bias / out_proj.*``, same maths).
PINNED: oracle/make_golden.py imports the reference file itself (with the three ``builtins`` names it relies on torch
leaking) and stores its output in tests/golden/resampler_small.pt.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn


def _sincos_1d(embed_dim, pos):
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("hw,d->hwd", pos, omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=-1)


def get_2d_sincos_pos_embed(embed_dim, image_size):
    """[H, W, embed_dim]: first half encodes the w coordinate grid, second half the h grid ("w goes first", :47)."""
    gh, gw = (image_size, image_size) if isinstance(image_size, int) else (image_size[0], image_size[1])
    grid = np.stack(np.meshgrid(np.arange(gw, dtype=np.float32), np.arange(gh, dtype=np.float32)), axis=0)
    return np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=-1)


class Resampler(nn.Module):
    def __init__(self, num_queries, embed_dim, num_heads, kv_dim=None, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                 adaptive=False, max_size=(70, 70)):
        super().__init__()
        self.num_queries, self.embed_dim, self.num_heads, self.max_size = num_queries, embed_dim, num_heads, max_size
        self.query = nn.Parameter(torch.zeros(num_queries, embed_dim))
        self.kv_proj = nn.Linear(kv_dim, embed_dim, bias=False) if kv_dim is not None and kv_dim != embed_dim else nn.Identity()
        self.attn = nn.MultiheadAttention(embed_dim, num_heads)
        self.ln_q, self.ln_kv, self.ln_post = norm_layer(embed_dim), norm_layer(embed_dim), norm_layer(embed_dim)
        self.proj = nn.Parameter((embed_dim ** -0.5) * torch.randn(embed_dim, embed_dim))
        self.register_buffer("pos_embed", torch.from_numpy(get_2d_sincos_pos_embed(embed_dim, max_size)).float(), persistent=False)

    def forward(self, x, tgt_sizes):
        bs = x.shape[0]
        mh, mw = int(tgt_sizes[:, 0].max()), int(tgt_sizes[:, 1].max())
        if mh > self.max_size[0] or mw > self.max_size[1]:  # _adjust_pos_cache (:130-135)
            self.max_size = [max(mh, self.max_size[0]), max(mw, self.max_size[1])]
            self.register_buffer("pos_embed", torch.from_numpy(get_2d_sincos_pos_embed(self.embed_dim, self.max_size)).float().to(x.device),
                                 persistent=False)
        patch_len = tgt_sizes[:, 0] * tgt_sizes[:, 1]
        max_len = int(patch_len.max())
        mask = torch.zeros(bs, max_len, dtype=torch.bool, device=x.device)
        pos = []
        for i in range(bs):
            h, w = int(tgt_sizes[i, 0]), int(tgt_sizes[i, 1])
            pos.append(self.pos_embed[:h, :w, :].reshape(h * w, -1).to(x.dtype))
            mask[i, int(patch_len[i]):] = True
        pos = torch.nn.utils.rnn.pad_sequence(pos, batch_first=True, padding_value=0.0).permute(1, 0, 2)
        x = self.ln_kv(self.kv_proj(x)).permute(1, 0, 2)
        q = self.ln_q(self.query)
        out = self.attn(q.unsqueeze(1).repeat(1, bs, 1), x + pos, x, key_padding_mask=mask)[0]
        return self.ln_post(out.permute(1, 0, 2)) @ self.proj
