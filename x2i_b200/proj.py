"""Drop-in alignment projector (``utils/proj.py`` of the reference) on sm_100a kernels.

Same classes, constructor arguments, parameter names and return values as the reference so that
``create_proj3_qwen3b(...)`` / ``proj.load_state_dict(...)`` / ``pooled, prompt_embeds = proj(text_embeddings)``
(``infer/inference_qwenvl.py:77-94,:179``; ``train/train_qwenvl.py:399-410,:576``) work unchanged:

    x[B,C,S,H] --(Conv2d(C->1,5x5) | cha_scale-mean | mean)--> [B,S,H] --LayerNorm--> Linear(H,4096,no bias)
      --GELU--> Linear(4096,4096,no bias) = prompt_embeds[B,S,4096] --GELU--> Linear(4096,768)+b --mean_S--> pooled[B,768]

Kernels: one HBM-streaming stencil+LayerNorm kernel for the front end (x2i_proj_mix_ln), tcgen05 GEMMs with GELU
epilogues for the three linears, a column-mean kernel.  Forward only for now (the projector's backward / wgrad is a
later scope row).  ``use_t5=True`` raises exactly like the reference (NameError there).
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16


class MLP3(nn.Module):
    def __init__(self, in_dim=4096, out_dim=4096, hidden_dim=4096, out_dim1=768, layer_norm_eps=1e-5, use_residual=True):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim, eps=layer_norm_eps)
        self.projector = nn.Sequential(nn.Linear(in_dim, hidden_dim, bias=False), nn.GELU(),
                                       nn.Linear(hidden_dim, hidden_dim, bias=False))
        self.fc = nn.Sequential(nn.GELU(), nn.Linear(out_dim, out_dim1))

    def _tail(self, xn):
        """xn: layer-normed [B,S,H] bf16."""
        h = ops.linear(xn, self.projector[0].weight, None, act=2)
        x2, g = ops.linear_dual_gelu(h, self.projector[2].weight, None)
        x1 = ops.mean_over_s(ops.linear(g, self.fc[1].weight, self.fc[1].bias))
        return x1, x2

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            raise X2IError("x2i_b200.proj: backward is not implemented yet; call under torch.no_grad()")
        B, S, H = x.shape
        one = torch.ones(1, device=x.device)
        xn = ops.proj_mix_ln(x.to(BF16).reshape(B, 1, S, H), 2, one, 0.0, self.layernorm.weight.float(),
                             self.layernorm.bias.float(), self.layernorm.eps)
        return self._tail(xn)


class Proj7Exp(nn.Module):
    def __init__(self, in_channels=25, kernel_size=5, input_dim=896, output_dim0=768, output_dim1=4096, num_layers=2,
                 num_heads=12, norm_eps=1e-6, head_dim=64, use_t5=True, use_scale=True, use_cnn=True) -> None:
        super().__init__()
        if use_t5:
            raise NameError("name 'T5Config' is not defined (the reference's use_t5=True branch is dead code: "
                            "utils/proj.py:42-46 never imports T5Config/T5Stack; every caller passes use_t5=False)")
        if kernel_size != 5:
            raise X2IError("Proj7Exp: the stencil kernel is specialised for kernel_size=5 (all reference factories)")
        self.use_t5, self.use_scale, self.use_cnn = use_t5, use_scale, use_cnn
        if self.use_scale:
            self.cha_scale = nn.Parameter(torch.empty(1, in_channels, 1, 1), requires_grad=True)
            nn.init.xavier_normal_(self.cha_scale, gain=1)
        elif self.use_cnn:
            self.conv = nn.Conv2d(in_channels, 1, kernel_size=kernel_size, padding=(kernel_size - 1) // 2)
        self.mlp = MLP3(input_dim, output_dim1, output_dim1, output_dim0, norm_eps)

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise X2IError("x2i_b200.proj: backward is not implemented yet; call under torch.no_grad()")
        if self.mlp.projector[0].weight.dtype != BF16 or not x.is_cuda:
            raise X2IError("Proj7Exp runs in bf16 on a CUDA device (x2i_b200 has no CPU path): .to('cuda', torch.bfloat16)")
        B, C, S, H = x.shape
        ln = self.mlp.layernorm
        gamma, beta = ln.weight.float(), ln.bias.float()
        xb = x.to(BF16)
        if self.use_scale:
            xn = ops.proj_mix_ln(xb, 1, self.cha_scale.float().reshape(C).contiguous(), 0.0, gamma, beta, ln.eps)
        elif self.use_cnn:
            xn = ops.proj_mix_ln(xb, 0, self.conv.weight.float().reshape(C, 25).contiguous(), float(self.conv.bias.float()),
                                 gamma, beta, ln.eps)
        else:
            xn = ops.proj_mix_ln(xb, 2, gamma, 0.0, gamma, beta, ln.eps)
        return self.mlp._tail(xn)


def _mk(in_channels, input_dim, num_heads, head_dim, use_t5, use_scale, use_cnn):
    return Proj7Exp(in_channels=in_channels, kernel_size=5, input_dim=input_dim, output_dim0=768, output_dim1=4096,
                    num_layers=2, num_heads=num_heads, norm_eps=1e-6, head_dim=head_dim, use_t5=use_t5, use_scale=use_scale,
                    use_cnn=use_cnn)


def create_proj3_qwen3b(in_channels, use_t5=True, use_scale=True, use_cnn=False):
    return _mk(in_channels, 2048, 28, 128, use_t5, use_scale, False if use_scale else use_cnn)


def create_proj3_qwen7b(in_channels, use_t5=True, use_scale=True, use_cnn=False):
    return _mk(in_channels, 3584, 28, 128, use_t5, use_scale, False if use_scale else use_cnn)


def create_proj_internvl1b(in_channels, use_t5=True, use_scale=True, use_cnn=True):
    return _mk(in_channels, 896, 12, 64, use_t5, use_scale, use_cnn)


def create_proj_internvl4b(in_channels, use_t5=True, use_scale=False, use_cnn=True):
    return _mk(in_channels, 2048, 16, 128, use_t5, use_scale, use_cnn)


def create_proj_minicpm(in_channels, use_t5=True, use_scale=True, use_cnn=False):
    return _mk(in_channels, 3584, 28, 128, use_t5, use_scale, False if use_scale else use_cnn)


def load_projector_state(module: nn.Module, state_dict) -> nn.Module:
    """Load a reference checkpoint (``diffusion_pytorch_model.bin``), stripping the DDP ``module.`` prefix as the
    reference loaders do (infer/inference_qwenvl.py:85-91)."""
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    module.load_state_dict(sd)
    return module
