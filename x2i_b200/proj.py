"""Drop-in alignment projector (``utils/proj.py`` of the reference) on sm_100a kernels.

Same classes, constructor arguments, parameter names and return values as the reference so that
``create_proj3_qwen3b(...)`` / ``proj.load_state_dict(...)`` / ``pooled, prompt_embeds = proj(text_embeddings)``
(``infer/inference_qwenvl.py:77-94,:179``; ``train/train_qwenvl.py:399-410,:576``) work unchanged:

    x[B,C,S,H] --(Conv2d(C->1,5x5) | cha_scale-mean | mean)--> [B,S,H] --LayerNorm--> Linear(H,4096,no bias)
      --GELU--> Linear(4096,4096,no bias) = prompt_embeds[B,S,4096] --GELU--> Linear(4096,768)+b --mean_S--> pooled[B,768]

Kernels: one HBM-streaming stencil+LayerNorm kernel for the front end (x2i_proj_mix_ln), tcgen05 GEMMs with GELU
epilogues for the three linears, a column-mean kernel.  With trainable parameters and grad enabled the module runs as one
autograd node (``_ProjFn``) whose backward is hand-written too: wgrad / dgrad tcgen05 GEMMs (MN-major operands, GELU'
epilogues), LayerNorm backward, a stencil weight-gradient kernel for the 5x5 layer-mixing conv, deterministic reductions.
``use_t5=True`` raises exactly like the reference (NameError there).
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16


class _ProjFn(torch.autograd.Function):
    """Projector forward + backward on the sm_100a kernels (the only trained module of the distillation step,
    train/train_qwenvl.py:453-459, :576, :625).  Inputs: x [B,C,S,H] and the parameters; outputs (pooled, seq).
    mode 0 conv, 1 cha_scale, 2 plain mean.  Gradients for every parameter (x is the frozen MLLM's output: no gradient)."""

    @staticmethod
    def forward(ctx, x, mode, mix_w, mix_b, ln_w, ln_b, eps, w0, w2, w3, b3):
        xb = x.detach().to(BF16).contiguous()
        B, C, S, H = xb.shape
        gamma, beta = ln_w.detach().float(), ln_b.detach().float()
        if mode == 0:
            wf, cb = mix_w.detach().float().reshape(C, 25).contiguous(), float(mix_b.detach().float())
        elif mode == 1:
            wf, cb = mix_w.detach().float().reshape(C).contiguous(), 0.0
        else:
            wf, cb = gamma, 0.0
        xn, xm = ops.proj_mix_ln_save(xb, mode, wf, cb, gamma, beta, eps)
        M = B * S
        h1, g1 = ops.linear_act_save(xn.view(M, H), w0.detach(), None, 2)
        x2, g2 = ops.linear_dual_gelu(g1, w2.detach(), None)
        x1 = ops.mean_over_s(ops.linear(g2, w3.detach(), b3.detach()).view(B, S, -1))
        ctx.save_for_backward(xb, xm, xn, h1, g1, x2, g2, ln_w, w0, w2, w3)
        ctx.meta = (mode, eps, B, C, S, H, None if mix_w is None else (mix_w.shape, mix_w.dtype), ln_w.dtype, w0.dtype, b3.dtype)
        return x1, x2.view(B, S, -1)

    @staticmethod
    def backward(ctx, dx1, dx2):
        xb, xm, xn, h1, g1, x2, g2, ln_w, w0, w2, w3 = ctx.saved_tensors
        mode, eps, B, C, S, H, mixmeta, ln_dt, w_dt, b_dt = ctx.meta
        M = B * S
        dev = xb.device
        dy = ops.mean_over_s_bwd(dx1.to(BF16), S).view(M, -1)                     # through mean over S
        db3 = torch.empty(1, dy.shape[1], device=dev, dtype=torch.float32)
        ops.colsum(dy, 1, M, out0=db3)
        dw3 = ops.linear_wgrad(dy, g2)
        dx2t = ops.linear_dgrad(dy, w3.detach(), pre=x2, n_split=0, dact=2, addend=dx2.to(BF16).contiguous().view(M, -1))
        dw2 = ops.linear_wgrad(dx2t, g1)
        dh1 = ops.linear_dgrad(dx2t, w2.detach(), pre=h1, n_split=0, dact=2)
        dw0 = ops.linear_wgrad(dh1, xn.view(M, H))
        dxn = ops.linear_dgrad(dh1, w0.detach())
        stats = torch.empty(M, 2, device=dev, dtype=torch.float32)
        gamma_b = ln_w.detach().to(BF16).reshape(1, H).contiguous()
        dxm = ops.ln_modulate_bwd(dxn, xm.view(M, H), gamma_b, M, stats=stats, eps=eps, affine=True)
        dbeta = torch.empty(1, H, device=dev, dtype=torch.float32)
        dgamma = torch.empty(1, H, device=dev, dtype=torch.float32)
        ops.colsum(dxn, 1, M, out0=dbeta, b=xm.view(M, H), out1=dgamma, stats=stats)
        d_mix_w = d_mix_b = None
        if mode in (0, 1):
            shape, dt = mixmeta
            d_mix_w = ops.proj_mix_wgrad(xb, dxm.view(B, S, H), mode).view(shape).to(dt)
            if mode == 0:
                tot = torch.empty(1, H, device=dev, dtype=torch.float32)
                ops.colsum(dxm, 1, M, out0=tot)
                d_mix_b = tot.sum().reshape(1).to(dt)
        return (None, None, d_mix_w, d_mix_b, dgamma.view(H).to(ln_dt), dbeta.view(H).to(ln_dt), None, dw0.to(w_dt), dw2.to(w_dt),
                dw3.to(w_dt), db3.view(-1).to(b_dt))


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
        if self.mlp.projector[0].weight.dtype != BF16 or not x.is_cuda:
            raise X2IError("Proj7Exp runs in bf16 on a CUDA device (x2i_b200 has no CPU path): .to('cuda', torch.bfloat16)")
        B, C, S, H = x.shape
        ln = self.mlp.layernorm
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if x.requires_grad:
                raise X2IError("Proj7Exp: no gradient towards the MLLM hidden states (the MLLM is frozen in X2I training)")
            m = self.mlp
            mode = 1 if self.use_scale else (0 if self.use_cnn else 2)
            mix_w = self.cha_scale if mode == 1 else (self.conv.weight if mode == 0 else None)
            mix_b = self.conv.bias if mode == 0 else None
            return _ProjFn.apply(x, mode, mix_w, mix_b, ln.weight, ln.bias, ln.eps, m.projector[0].weight, m.projector[2].weight,
                                 m.fc[1].weight, m.fc[1].bias)
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


# ---- ComfyUI bundle format (x2i_comfyui/model.py:33-39, :89-97): {"config": ctor kwargs, "state_dict": ...} in one .pt -------
def load_projector_bundle(path: str, device="cuda") -> Proj7Exp:
    """``Proj.load(path)`` of the ComfyUI nodes: build the projector from the bundle's config and load its weights."""
    all_dict = torch.load(path, map_location="cpu", weights_only=True)
    proj = Proj7Exp(**all_dict["config"])
    load_projector_state(proj, all_dict["state_dict"])
    return proj.eval().to(device, BF16) if device is not None else proj.eval()


def save_projector_bundle(proj: Proj7Exp, path: str, config: dict | None = None) -> None:
    """``Proj.transfrom(config, state_path, save_path)``: write {"config", "state_dict"}; config defaults to the module's own."""
    if config is None:
        m = proj.mlp
        config = dict(in_channels=(proj.cha_scale.shape[1] if proj.use_scale else proj.conv.in_channels if proj.use_cnn else 1),
                      kernel_size=5, input_dim=m.layernorm.normalized_shape[0], output_dim0=m.fc[1].out_features,
                      output_dim1=m.projector[2].out_features, num_layers=2, num_heads=16, norm_eps=m.layernorm.eps, head_dim=128,
                      use_t5=False, use_scale=proj.use_scale, use_cnn=proj.use_cnn)
    torch.save({"config": config, "state_dict": {k: v.detach().cpu() for k, v in proj.state_dict().items()}}, path)
