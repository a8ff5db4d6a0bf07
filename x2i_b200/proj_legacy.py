"""The older alignment-projector variants of ``model_internvl/proj.py`` on the x2i_b200 kernels, inference only.

Reference classes (``/root/reference/model_internvl/proj.py``; imported by nothing in the reference tree, kept for its checkpoints):
  ``MLP`` :53-74, ``MLP2`` :77-104, ``MLP_plus`` :106-134  -- LayerNorm -> 3 / 3 / 6 bias-free Linear with GELU(erf) between -> x;
      x2 = GELU(x); x1 = mean_S(fc(x2)); returns (x1, x2)   (note: x2 is POST-GELU here, unlike utils/proj.py's MLP3);
  ``Proj`` :155-173, ``Proj2`` :175-193  -- LayerNorm(H) -> Conv2d(C -> 1, 5x5) over (S, H) -> LayerNorm -> T5Stack -> MLP / MLP2;
  ``Proj3`` :196-211                      -- T5Stack on every layer's sequence first, then LayerNorm -> conv -> LayerNorm -> MLP2.
``Transformer_proj`` (:137-153, an nn.TransformerEncoder variant) is not provided.

The T5Stack is transformers' encoder (``T5Config(feed_forward_proj="gated-gelu", dense_act_fn="gelu_new", is_decoder=False)``,
proj.py:158-160): per block T5LayerNorm (RMS) -> bias-free q/k/v -> scores = q k^T (NO 1/sqrt(d) scaling) + bucketed relative position
bias -> softmax -> o + residual; T5LayerNorm -> wo(gelu_new(wi_0 x) * wi_1 x) + residual; final T5LayerNorm.  Kernels: x2i_rmsnorm, the
tcgen05 GEMMs (fused q|k|v, GEGLU epilogue, residual epilogues), and -- because head_dim is 64 and the scores carry an additive bias,
which the fused d = 128 attention kernel does not do -- the explicit attention path of the VAE mid block: fp32 score GEMM ->
x2i_softmax_rows_bias -> P.V GEMM, one (sequence, head) at a time.  That is slow (dead code in the reference: correctness first).
Sequence length must be a multiple of 32.  No CPU fallback; no backward (use utils/proj.py's Proj7Exp family for training).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16


def _need_cuda_bf16(w):
    if w.dtype != BF16 or not w.is_cuda:
        raise X2IError("x2i_b200.proj_legacy modules run in bf16 on a CUDA device: call .to('cuda', torch.bfloat16) first")


def _no_grad(x):
    if torch.is_grad_enabled() and x.requires_grad:
        raise X2IError("x2i_b200.proj_legacy is inference only (the reference never trains these variants); use torch.no_grad()")


class _MLPBase(nn.Module):
    """Shared forward of MLP / MLP2 / MLP_plus: ``projector`` = bias-free Linears with GELU(erf) between, ``fc`` = head on GELU(x)."""

    def _run(self, x):
        _need_cuda_bf16(self.layernorm.weight)
        _no_grad(x)
        B, S, H = x.shape
        h = ops.layernorm_affine(x.to(BF16).reshape(B * S, H), self.layernorm.weight, self.layernorm.bias, self.layernorm.eps)
        lins = [m for m in self.projector if isinstance(m, nn.Linear)]
        for lin in lins[:-1]:
            h = ops.linear(h, lin.weight, None, act=2)
        _, x2 = ops.linear_dual_gelu(h, lins[-1].weight, None)  # (x, GELU(x)): the reference returns the activated tensor
        if isinstance(self.fc, nn.Linear):
            y = ops.linear(x2, self.fc.weight, self.fc.bias)
        else:
            fcs = [m for m in self.fc if isinstance(m, nn.Linear)]
            y = x2
            for i, lin in enumerate(fcs):
                y = ops.linear(y, lin.weight, None, act=2 if i + 1 < len(fcs) else 0)
        x1 = ops.mean_over_s(y.view(B, S, -1))
        return x1, x2.view(B, S, -1)

    def forward(self, x):
        return self._run(x)


def _seq(dims, bias=False):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(nn.Linear(dims[i], dims[i + 1], bias=bias))
        if i + 2 < len(dims):
            mods.append(nn.GELU())
    return nn.Sequential(*mods)


class MLP(_MLPBase):
    def __init__(self, in_dim=4096, out_dim=4096, hidden_dim=4096, out_dim1=768, layer_norm_eps=1e-5, use_residual=True):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim, eps=layer_norm_eps)
        self.projector = _seq([in_dim, hidden_dim, hidden_dim, out_dim])
        self.fc = nn.Linear(out_dim, out_dim1)


class MLP2(_MLPBase):
    def __init__(self, in_dim=4096, out_dim=4096, hidden_dim=4096, out_dim1=768, layer_norm_eps=1e-5, use_residual=True):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim, eps=layer_norm_eps)
        self.projector = _seq([in_dim, hidden_dim, hidden_dim, out_dim])
        self.fc = _seq([out_dim, out_dim1, out_dim1, out_dim1])


class MLP_plus(_MLPBase):
    def __init__(self, in_dim=4096, out_dim=4096, hidden_dim=4096, out_dim1=768, use_residual=True):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim)
        self.projector = _seq([in_dim] + [hidden_dim] * 5 + [out_dim])
        self.fc = nn.Linear(out_dim, out_dim1)


# ------------------------------------------------------------------------------------------------ T5 encoder stack
class _T5Norm(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _T5Attention(nn.Module):
    def __init__(self, d_model, heads, d_kv, has_bias, num_buckets):
        super().__init__()
        inner = heads * d_kv
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(num_buckets, heads)


class _T5SelfAttentionLayer(nn.Module):
    def __init__(self, d_model, heads, d_kv, eps, has_bias, num_buckets):
        super().__init__()
        self.SelfAttention = _T5Attention(d_model, heads, d_kv, has_bias, num_buckets)
        self.layer_norm = _T5Norm(d_model, eps)


class _T5Dense(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi_0 = nn.Linear(d_model, d_ff, bias=False)
        self.wi_1 = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)


class _T5FFLayer(nn.Module):
    def __init__(self, d_model, d_ff, eps):
        super().__init__()
        self.DenseReluDense = _T5Dense(d_model, d_ff)
        self.layer_norm = _T5Norm(d_model, eps)


class _T5Block(nn.Module):
    def __init__(self, d_model, heads, d_kv, d_ff, eps, has_bias, num_buckets):
        super().__init__()
        self.layer = nn.ModuleList([_T5SelfAttentionLayer(d_model, heads, d_kv, eps, has_bias, num_buckets), _T5FFLayer(d_model, d_ff, eps)])


def relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """T5Attention._relative_position_bucket, bidirectional (encoder).  Integer index work: bit-exact."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    rp = relative_position.abs()
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


class T5Stack(nn.Module):
    """Encoder-only ``transformers`` T5Stack as configured at model_internvl/proj.py:158-160; parameter names follow the library
    (``block.N.layer.0.SelfAttention.q.weight`` ...).  The library's unused ``embed_tokens`` table is not held (dropped on load)."""

    def __init__(self, d_model, num_layers, num_heads, d_kv, d_ff, eps=1e-6, num_buckets=32, max_distance=128):
        super().__init__()
        if d_kv % 64 or d_model % 32 or d_ff % 128 or (num_heads * d_kv) % 32:
            raise X2IError("T5Stack (x2i_b200): d_kv % 64, d_model % 32, d_ff % 128 must be 0")
        self.cfg = dict(d_model=d_model, num_layers=num_layers, num_heads=num_heads, d_kv=d_kv, d_ff=d_ff, eps=eps, num_buckets=num_buckets,
                        max_distance=max_distance)
        self.block = nn.ModuleList([_T5Block(d_model, num_heads, d_kv, d_ff, eps, i == 0, num_buckets) for i in range(num_layers)])
        self.final_layer_norm = _T5Norm(d_model, eps)
        self._packed = None

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        state_dict.pop(prefix + "embed_tokens.weight", None)  # unused by inputs_embeds=...; 32128 x d_model in the library's module
        return super()._load_from_state_dict(state_dict, prefix, *a, **k)

    def _pack(self):
        w0 = self.block[0].layer[0].SelfAttention.q.weight
        _need_cuda_bf16(w0)
        if self._packed is None or self._packed[0] != (w0.data_ptr(), w0._version):
            packs = []
            for b in self.block:
                a, f = b.layer[0].SelfAttention, b.layer[1].DenseReluDense
                packs.append(dict(w_qkv=torch.cat([a.q.weight, a.k.weight, a.v.weight], 0).contiguous(),
                                  w_gu=ops.pack_swiglu_weight(f.wi_0.weight.detach(), f.wi_1.weight.detach())))
            self._packed = ((w0.data_ptr(), w0._version), packs)
        return self._packed[1]

    def position_bias(self, S, device):
        """[heads, S, S] fp32: relative_attention_bias(bucket(j - i)) (T5Attention.compute_bias); shared by all layers and sequences."""
        c = self.cfg
        ctx = torch.arange(S, device=device)[:, None]
        mem = torch.arange(S, device=device)[None, :]
        bucket = relative_position_bucket(mem - ctx, c["num_buckets"], c["max_distance"])
        emb = self.block[0].layer[0].SelfAttention.relative_attention_bias.weight  # [buckets, heads]
        return emb[bucket].permute(2, 0, 1).float().contiguous()

    @torch.no_grad()
    def forward(self, inputs_embeds=None, **unused):
        if inputs_embeds is None:
            raise X2IError("T5Stack (x2i_b200): call with inputs_embeds=..., as model_internvl/proj.py does")
        c = self.cfg
        packs = self._pack()
        x = inputs_embeds.to(BF16).contiguous()
        N, S, D = x.shape
        if S % 32:
            raise X2IError("T5Stack (x2i_b200): the explicit attention path needs a sequence length that is a multiple of 32")
        heads, dk = c["num_heads"], c["d_kv"]
        inner = heads * dk
        dev = x.device
        bias = self.position_bias(S, dev)
        h = x.clone()
        h2 = torch.empty_like(h)
        xn = torch.empty(N, S, D, device=dev, dtype=BF16)
        qkv = torch.empty(N * S, 3 * inner, device=dev, dtype=BF16)
        att = torch.empty(N * S, inner, device=dev, dtype=BF16)
        act = torch.empty(N * S, c["d_ff"], device=dev, dtype=BF16)
        scores = torch.empty(S, S, device=dev, dtype=torch.float32)
        P = torch.empty(S, S, device=dev, dtype=BF16)
        for blk, pk in zip(self.block, packs):
            sa, ff = blk.layer[0], blk.layer[1]
            ops.rmsnorm(h, sa.layer_norm.weight, c["eps"], out=xn)
            ops.linear(xn.view(N * S, D), pk["w_qkv"], None, out=qkv)
            for n in range(N):
                rows = slice(n * S, (n + 1) * S)
                for hh in range(heads):
                    q = qkv[rows, hh * dk:(hh + 1) * dk]
                    k = qkv[rows, inner + hh * dk:inner + (hh + 1) * dk]
                    v = qkv[rows, 2 * inner + hh * dk:2 * inner + (hh + 1) * dk]
                    ops.linear_f32(q, k, 1.0, out=scores)                          # T5: no 1/sqrt(d) scaling
                    ops.softmax_rows_bias(scores, bias[hh], out=P)
                    ops.linear_dgrad(P, v, out=att[rows, hh * dk:(hh + 1) * dk])      # P @ v: v is the [K, N] N-contiguous operand
            ops.linear_residual(att, sa.SelfAttention.o.weight, h.view(N * S, D), h2.view(N * S, D))
            ops.rmsnorm(h2, ff.layer_norm.weight, c["eps"], out=xn)
            ops.linear_swiglu(xn.view(N * S, D), pk["w_gu"], out=act, act=1)      # gelu_new(wi_0 x) * wi_1 x
            ops.linear_residual(act, ff.DenseReluDense.wo.weight, h2.view(N * S, D), h.view(N * S, D))
        out = ops.rmsnorm(h, self.final_layer_norm.weight, c["eps"])
        return _LastHidden(out)


class _LastHidden:
    def __init__(self, t):
        self.last_hidden_state = t

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class _ProjBase(nn.Module):
    mlp_cls = MLP
    t5_first = False

    def __init__(self, in_channels=2, kernel_size=5, input_dim=896, output_dim0=768, output_dim1=4096, num_layers=4, num_heads=12,
                 layer_norm_eps=1e-6, head_dim=64):
        super().__init__()
        if kernel_size != 5:
            raise X2IError("proj_legacy: the layer-mixing stencil kernel is 5x5 (every reference configuration)")
        self.norm0 = nn.LayerNorm(input_dim, eps=layer_norm_eps)
        self.conv = nn.Conv2d(in_channels, 1, kernel_size=kernel_size, padding=(kernel_size - 1) // 2)
        self.norm1 = nn.LayerNorm(input_dim, eps=layer_norm_eps)
        self.t5stack = T5Stack(input_dim, num_layers, num_heads, head_dim, input_dim * 4, layer_norm_eps)
        self.mlp = self.mlp_cls(input_dim, output_dim1, output_dim1, output_dim0, layer_norm_eps)

    def _front(self, x):
        """norm0 -> conv(C -> 1, 5x5) -> squeeze -> norm1   ([B, C, S, H] -> [B, S, H])."""
        B, C, S, H = x.shape
        xn = ops.layernorm_affine(x.to(BF16).reshape(B * C * S, H), self.norm0.weight, self.norm0.bias, self.norm0.eps).view(B, C, S, H)
        return ops.proj_mix_ln(xn, 0, self.conv.weight.float().reshape(C, 25).contiguous(), float(self.conv.bias.float()),
                               self.norm1.weight.float(), self.norm1.bias.float(), self.norm1.eps)

    @torch.no_grad()
    def forward(self, x):
        _need_cuda_bf16(self.norm0.weight)
        if self.t5_first:
            B, C, S, H = x.shape
            x = self.t5stack(inputs_embeds=x.contiguous().view(B * C, S, H)).last_hidden_state.view(B, C, S, H)
            return self.mlp(self._front(x))
        return self.mlp(self.t5stack(inputs_embeds=self._front(x)).last_hidden_state)


class Proj(_ProjBase):
    mlp_cls = MLP


class Proj2(_ProjBase):
    mlp_cls = MLP2


class Proj3(_ProjBase):
    mlp_cls = MLP2
    t5_first = True
