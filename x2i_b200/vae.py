"""Drop-in ``AutoencoderKL`` (decoder for the step right after the denoise loop; plus the encoder) on sm_100a kernels.

Reference call sites: ``infer/inference_qwenvl.py:75`` (``AutoencoderKL.from_pretrained(flux_path, subfolder="vae",
torch_dtype=dtype).to(device)``) and ``:209-216`` (``vae.config.block_out_channels / scaling_factor / shift_factor``,
``vae.decode(latents, return_dict=False)[0]``, ``VaeImageProcessor.postprocess``).  The class lives in diffusers 0.31.0
[D031]; parameter names / state-dict keys are diffusers' (``decoder.conv_in``, ``decoder.mid_block.resnets.N``,
``decoder.mid_block.attentions.0.{group_norm,to_q,to_k,to_v,to_out.0}``, ``decoder.up_blocks.N.resnets.M``,
``decoder.up_blocks.N.upsamplers.0.conv``, ``decoder.conv_norm_out``, ``decoder.conv_out``) so the published checkpoint loads.

Inside, activations are NHWC bf16.  Every convolution is the implicit-GEMM tcgen05 kernel of the ControlNeXt branch
(``x2i_conv2d_nhwc``; the 16-channel input and the 3-channel output are zero-padded to 64 channels, the residual add of a
ResnetBlock2D is fused into its second conv's epilogue), GroupNorm(32) + SiLU is the fused deterministic kernel pair,
nearest 2x upsampling is one copy kernel, and the single-head d=512 mid-block attention over the 128x128 latent pixels is
three tcgen05 GEMMs around a row soft-max with fp32 scores (``x2i_gemm_f32`` -> ``x2i_softmax_rows`` -> ``x2i_gemm_kn``).
``encode`` (``vae.encode(pixel_values).latent_dist.sample()``, ``lightcontrol/train_lightcontrol.py:678``) runs the encoder half on
the same kernels; its down-sampling convs (``F.pad(x, (0, 1, 0, 1))`` + stride-2 conv) use the conv kernel's separate trailing
padding.  Inference only (the VAE is frozen everywhere in the reference).  No CPU / eager fallback.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16

FLUX_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, act_fn="silu", scaling_factor=0.3611, shift_factor=0.1159, use_quant_conv=False,
                       use_post_quant_conv=False, mid_block_add_attention=True, force_upcast=True, sample_size=1024)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise X2IError(f"{type(self).__name__} is a parameter holder inside the fused AutoencoderKL decoder")


class ResnetBlock2D(_Holder):
    def __init__(self, in_channels, out_channels, groups, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Attention(_Holder):
    def __init__(self, channels, groups, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])


class UNetMidBlock2D(_Holder):
    def __init__(self, channels, groups):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(channels, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, groups), ResnetBlock2D(channels, channels, groups)])


class Upsample2D(_Holder):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class UpDecoderBlock2D(_Holder):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class Downsample2D(_Holder):
    """Conv2d(k3, s2, p0) applied after F.pad(x, (0, 1, 0, 1)) -- the padding is the conv kernel's pad_end."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)


class DownEncoderBlock2D(_Holder):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class Encoder(_Holder):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        prev = block_out_channels[0]
        for i, ch in enumerate(block_out_channels):
            self.down_blocks.append(DownEncoderBlock2D(prev, ch, layers_per_block, i != len(block_out_channels) - 1, norm_num_groups))
            prev = ch
        self.mid_block = UNetMidBlock2D(prev, norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, prev, eps=1e-6)
        self.conv_out = nn.Conv2d(prev, 2 * latent_channels, 3, padding=1)


class DiagonalGaussianDistribution:
    """``vae.encode(x).latent_dist`` (diffusers [D031]): mean | logvar = chunk(2, dim=1), logvar clamped to [-30, 20]."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class Decoder(_Holder):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], norm_num_groups)
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, ch in enumerate(rev):
            self.up_blocks.append(UpDecoderBlock2D(prev, ch, layers_per_block + 1, i != len(rev) - 1, norm_num_groups))
            prev = ch
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)


class AutoencoderKL(nn.Module):
    """``vae.decode(z, return_dict=False)[0]``, ``vae.encode(x).latent_dist`` and ``vae.config`` of the reference's VAE."""

    def __init__(self, **config):
        super().__init__()
        cfg = dict(FLUX_VAE_CONFIG)
        cfg.update(config)
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        if cfg.get("use_post_quant_conv"):
            raise X2IError("AutoencoderKL: use_post_quant_conv=True is not the FLUX configuration")
        if any(c % 64 for c in cfg["block_out_channels"]):
            raise X2IError("AutoencoderKL: block_out_channels must be multiples of 64")
        self.config = SimpleNamespace(**cfg)
        self.decoder = Decoder(cfg["latent_channels"], cfg["out_channels"], cfg["block_out_channels"], cfg["layers_per_block"],
                               cfg["norm_num_groups"])
        self.encoder = Encoder(cfg["in_channels"], cfg["latent_channels"], cfg["block_out_channels"], cfg["layers_per_block"],
                               cfg["norm_num_groups"])
        self._packed = {}

    @property
    def dtype(self):
        return self.decoder.conv_in.weight.dtype

    @property
    def device(self):
        return self.decoder.conv_in.weight.device

    # ------------------------------------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, **_kw):
        """Load a diffusers VAE directory (config.json + diffusion_pytorch_model.safetensors / .bin); quant-conv weights, if the
        file has any, are ignored (the FLUX configuration has none)."""
        import glob
        import json
        import os
        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, "config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_") and k in FLUX_VAE_CONFIG}
        m = cls(**cfg)
        sd = {}
        files = sorted(glob.glob(os.path.join(d, "*.safetensors")))
        if files:
            from safetensors.torch import load_file
            for fn in files:
                sd.update(load_file(fn))
        else:
            for fn in sorted(glob.glob(os.path.join(d, "*.bin"))):
                sd.update(torch.load(fn, map_location="cpu"))
        sd = {k: v for k, v in sd.items() if k.startswith(("decoder.", "encoder."))}
        m.load_state_dict(sd, strict=True)
        return m.to(torch_dtype) if torch_dtype is not None else m

    # ------------------------------------------------------------------------------------------------ weight packing
    def _w(self, conv: nn.Conv2d, pad_in=0, pad_out=0):
        """bf16 [Cout(+pad), kh*kw*(Cin+pad)] B operand of the implicit-GEMM conv, packed once per parameter version."""
        key = id(conv)
        ver = (conv.weight.data_ptr(), conv.weight._version, conv.bias._version)
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            w, b = conv.weight.detach(), conv.bias.detach()
            if pad_in:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_in, *w.shape[2:])], 1)
            if pad_out:
                w = torch.cat([w, w.new_zeros(pad_out, *w.shape[1:])], 0)
                b = torch.cat([b, b.new_zeros(pad_out)])
            hit = (ver, ops.pack_conv_weight(w), b.to(BF16).contiguous())
            self._packed[key] = hit
        return hit[1], hit[2]

    def _conv(self, x, conv: nn.Conv2d, residual=None, pad_in=0, pad_out=0, pad_end=None):
        w, b = self._w(conv, pad_in, pad_out)
        return ops.conv2d_nhwc(x, w, b, conv.kernel_size[0], conv.kernel_size[1], stride=conv.stride[0], pad=conv.padding[0],
                               residual=residual, pad_end=pad_end)

    @staticmethod
    def _gn(x, gn: nn.GroupNorm, act):
        return ops.groupnorm_nhwc(x, gn.weight, gn.bias, gn.num_groups, gn.eps, act=act)

    def _resnet(self, x, r: ResnetBlock2D):
        h = self._conv(self._gn(x, r.norm1, 2), r.conv1)
        skip = x if r.conv_shortcut is None else self._conv(x, r.conv_shortcut)
        return self._conv(self._gn(h, r.norm2, 2), r.conv2, residual=skip)       # (x + h) / output_scale_factor(1)

    def _attention(self, x, a: Attention):
        """AttnProcessor2_0 on the flattened pixels: NHWC rows ARE the tokens, so nothing is transposed."""
        B, H, W, C = x.shape
        T = H * W
        h = self._gn(x, a.group_norm, 0).view(B * T, C)
        q = ops.linear(h, a.to_q.weight, a.to_q.bias)
        k = ops.linear(h, a.to_k.weight, a.to_k.bias)
        v = ops.linear(h, a.to_v.weight, a.to_v.bias)
        o = torch.empty(B * T, C, device=x.device, dtype=BF16)
        scores = torch.empty(T, T, device=x.device, dtype=torch.float32)
        probs = torch.empty(T, T, device=x.device, dtype=BF16)
        for b in range(B):  # one image at a time: the fp32 score matrix is T*T*4 bytes (1 GiB at 1024 px)
            sl = slice(b * T, (b + 1) * T)
            ops.linear_f32(q[sl], k[sl], alpha=1.0 / math.sqrt(C), out=scores)
            ops.softmax_rows(scores, out=probs)
            o[sl] = ops.matmul_kn(probs, v[sl])
        out = ops.conv2d_nhwc(o.view(B, H, W, C), a.to_out[0].weight.detach().view(C, C).contiguous(), a.to_out[0].bias, 1, 1, stride=1,
                              pad=0, residual=x)                                 # to_out[0](o) + residual
        return out

    # ------------------------------------------------------------------------------------------------ decode
    def decode(self, z, return_dict=False, generator=None):
        d = self.decoder
        if d.conv_in.weight.dtype != BF16 or not z.is_cuda:
            raise X2IError("AutoencoderKL runs in bf16 on a CUDA device (x2i_b200 has no CPU path): .to('cuda', torch.bfloat16)")
        if torch.is_grad_enabled() and z.requires_grad:
            raise X2IError("AutoencoderKL.decode: inference only; call under torch.no_grad()")
        B, C, H, W = z.shape
        if C != self.config.latent_channels:
            raise X2IError(f"AutoencoderKL.decode: expected {self.config.latent_channels} latent channels, got {C}")
        cin_pad = (-C) % 64
        x = z.to(BF16).permute(0, 2, 3, 1)
        x = torch.cat([x, x.new_zeros(B, H, W, cin_pad)], -1).contiguous() if cin_pad else x.contiguous()
        x = self._conv(x, d.conv_in, pad_in=cin_pad)
        mid = d.mid_block
        x = self._resnet(x, mid.resnets[0])
        x = self._attention(x, mid.attentions[0])
        x = self._resnet(x, mid.resnets[1])
        for blk in d.up_blocks:
            for r in blk.resnets:
                x = self._resnet(x, r)
            if blk.upsamplers is not None:
                x = self._conv(ops.upsample2x_nhwc(x), blk.upsamplers[0].conv)
        x = self._gn(x, d.conv_norm_out, 2)
        cout = self.config.out_channels
        y = self._conv(x, d.conv_out, pad_out=(-cout) % 64)
        img = y[..., :cout].permute(0, 3, 1, 2).contiguous()                     # NCHW like the reference
        if return_dict:
            return SimpleNamespace(sample=img)
        return (img,)

    # ------------------------------------------------------------------------------------------------ encode
    def encode(self, x, return_dict=True):
        """``vae.encode(pixel_values).latent_dist.sample()`` (lightcontrol/train_lightcontrol.py:678): x [B, 3, H, W] in [-1, 1]
        -> DiagonalGaussianDistribution over [B, latent_channels, H/8, W/8]."""
        e = self.encoder
        if e.conv_in.weight.dtype != BF16 or not x.is_cuda:
            raise X2IError("AutoencoderKL runs in bf16 on a CUDA device (x2i_b200 has no CPU path): .to('cuda', torch.bfloat16)")
        if torch.is_grad_enabled() and x.requires_grad:
            raise X2IError("AutoencoderKL.encode: inference only (the VAE is frozen in the reference); call under torch.no_grad()")
        B, C, H, W = x.shape
        nd = len(e.down_blocks) - 1
        if C != self.config.in_channels or H % (1 << nd) or W % (1 << nd):
            raise X2IError(f"AutoencoderKL.encode: expected [B, {self.config.in_channels}, H, W] with H, W multiples of {1 << nd}")
        cin_pad = (-C) % 64
        h = x.to(BF16).permute(0, 2, 3, 1)
        h = torch.cat([h, h.new_zeros(B, H, W, cin_pad)], -1).contiguous() if cin_pad else h.contiguous()
        h = self._conv(h, e.conv_in, pad_in=cin_pad)
        for blk in e.down_blocks:
            for r in blk.resnets:
                h = self._resnet(h, r)
            if blk.downsamplers is not None:
                h = self._conv(h, blk.downsamplers[0].conv, pad_end=1)       # F.pad(x, (0, 1, 0, 1)) + Conv2d(k3, s2, p0)
        mid = e.mid_block
        h = self._resnet(h, mid.resnets[0])
        h = self._attention(h, mid.attentions[0])
        h = self._resnet(h, mid.resnets[1])
        h = self._gn(h, e.conv_norm_out, 2)
        cout = 2 * self.config.latent_channels
        y = self._conv(h, e.conv_out, pad_out=(-cout) % 64)
        dist = DiagonalGaussianDistribution(y[..., :cout].permute(0, 3, 1, 2).contiguous())
        if return_dict:
            return SimpleNamespace(latent_dist=dist)
        return (dist,)

    def forward(self, z):
        return self.decode(z)[0]


class VaeImageProcessor:
    """The two members of diffusers' VaeImageProcessor that infer/inference_qwenvl.py:210,:216 uses."""

    def __init__(self, vae_scale_factor=8, **_kw):
        self.vae_scale_factor = vae_scale_factor

    @staticmethod
    def denormalize(images):
        return (images / 2 + 0.5).clamp(0, 1)

    def postprocess(self, image, output_type="pil"):
        img = self.denormalize(image.float())
        if output_type in ("pt", "latent"):
            return img
        arr = img.cpu().permute(0, 2, 3, 1).numpy()
        if output_type == "np":
            return arr
        if output_type == "pil":
            from PIL import Image
            return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
        raise X2IError(f"VaeImageProcessor.postprocess: unknown output_type {output_type!r}")


def decode_latents(vae: AutoencoderKL, packed_latents, height, width):
    """infer/inference_qwenvl.py:209-216 in one call: packed [B, L, 64] latents -> [B, 3, height, width] image in [0, 1]."""
    from .pipeline import FluxPipeline
    scale = 2 ** len(vae.config.block_out_channels)
    z = FluxPipeline._unpack_latents(packed_latents, height, width, scale)
    z = (z / vae.config.scaling_factor) + vae.config.shift_factor
    return VaeImageProcessor.denormalize(vae.decode(z)[0])
