"""ORACLE (test infrastructure only) for the VAE decode that follows the denoise loop.

Reference call site: ``infer/inference_qwenvl.py:75`` (``AutoencoderKL.from_pretrained(flux_path, subfolder="vae")``) and
``:209-216``::

    latents = FluxPipeline._unpack_latents(latents, height, width, vae_scale_factor)      # 2 ** len(block_out_channels) = 16
    latents = (latents / vae.config.scaling_factor) + vae.config.shift_factor
    image = vae.decode(latents, return_dict=False)[0]
    image = image_processor.postprocess(image, output_type="pil")                        # (x / 2 + 0.5).clamp(0, 1)

``AutoencoderKL`` lives in ``diffusers==0.31.0`` (``models/autoencoders/autoencoder_kl.py``, ``vae.py``, ``unets/unet_2d_blocks.py``,
``resnet.py``, ``upsampling.py``, ``attention_processor.py``; un-vendored, requirements.txt:3).  This file restates the DECODER (and, for
``vae.encode`` of ``lightcontrol/train_lightcontrol.py:678``, the ENCODER) from the published architecture for the FLUX configuration (latent_channels 16, block_out_channels (128, 256, 512, 512),
layers_per_block 2, norm_num_groups 32, SiLU, mid-block attention with one 512-wide head, no post_quant_conv,
scaling_factor 0.3611, shift_factor 0.1159) with diffusers' state-dict key names.

PARITY UNPINNED by the reference (no tests / fixtures; diffusers not installable here).  Sanity anchor: the BFL-derived
autoencoder that ships in this image (``torchtitan.experiments.flux.model.autoencoder``, the implementation the diffusers
checkpoint was converted from) agrees with this restatement to fp32 round-off under the weight remapping in
``tests/test_oracle_golden.py::test_vae_oracle_matches_bfl_decoder``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

FLUX_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                       norm_num_groups=32, act_fn="silu", scaling_factor=0.3611, shift_factor=0.1159, use_quant_conv=False,
                       use_post_quant_conv=False, mid_block_add_attention=True, force_upcast=True, sample_size=1024)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps 1e-6, SiLU, output_scale_factor 1 [D031]."""

    def __init__(self, in_channels, out_channels, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers Attention as built by UNetMidBlock2D for the VAE: heads = 1, dim_head = channels, GroupNorm(32) first,
    bias everywhere, residual connection, rescale_output_factor 1; AttnProcessor2_0 on the flattened pixels [D031]."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(B, C, H, W) + x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, groups=32):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(channels, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, groups), ResnetBlock2D(channels, channels, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, in_channels=16, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_num_groups=32):
        super().__init__()
        rev = list(reversed(block_out_channels))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0], norm_num_groups)
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, ch in enumerate(rev):
            self.up_blocks.append(UpDecoderBlock2D(prev, ch, layers_per_block + 1, add_upsample=i != len(rev) - 1, groups=norm_num_groups))
            prev = ch
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, block_out_channels[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=0) of the VAE encoder: F.pad(x, (0, 1, 0, 1)) then Conv2d(k3, s2, p0) [D031]."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, groups) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class Encoder(nn.Module):
    """diffusers Encoder(double_z=True): conv_in, 4 DownEncoderBlock2D (2 resnets each, down-sampling on all but the last),
    the same mid block as the decoder, GroupNorm + SiLU + conv_out to 2 * latent_channels (mean | logvar) [D031]."""

    def __init__(self, in_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_num_groups=32):
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

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    """diffusers DiagonalGaussianDistribution: mean | logvar = chunk(2, dim=1), logvar clamped to [-30, 20] [D031]."""

    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class AutoencoderKLDecoder(nn.Module):
    """``vae.decode(z, return_dict=False)[0]`` of the FLUX AutoencoderKL (use_post_quant_conv=False: the decoder is applied
    to z directly).  Keys: ``decoder.*`` as in the diffusers checkpoint."""

    def __init__(self, **config):
        super().__init__()
        cfg = dict(FLUX_VAE_CONFIG)
        cfg.update(config)
        self.cfg = cfg
        self.decoder = Decoder(cfg["latent_channels"], cfg["out_channels"], tuple(cfg["block_out_channels"]), cfg["layers_per_block"],
                               cfg["norm_num_groups"])

    def decode(self, z, return_dict=False):
        return (self.decoder(z),)


class AutoencoderKL(AutoencoderKLDecoder):
    """Encoder + decoder: ``vae.encode(pixel_values).latent_dist.sample()`` of lightcontrol/train_lightcontrol.py:678 (no
    quant_conv in the FLUX configuration) and ``vae.decode``.  Keys ``encoder.*`` / ``decoder.*`` as in the checkpoint."""

    def __init__(self, **config):
        super().__init__(**config)
        cfg = self.cfg
        self.encoder = Encoder(cfg["in_channels"], cfg["latent_channels"], tuple(cfg["block_out_channels"]), cfg["layers_per_block"],
                               cfg["norm_num_groups"])

    def encode(self, x):
        from types import SimpleNamespace
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.encoder(x)))


def decode_latents(vae, packed_latents, height, width):
    """infer/inference_qwenvl.py:209-216 up to (and including) the [0, 1] image tensor of postprocess()."""
    from .flux_oracle import unpack_latents
    scale = 2 ** len(vae.cfg["block_out_channels"])
    z = unpack_latents(packed_latents, height, width, scale)
    z = z / vae.cfg["scaling_factor"] + vae.cfg["shift_factor"]
    img = vae.decode(z)[0]
    return (img / 2 + 0.5).clamp(0, 1)


def bfl_encoder_key_map(cfg=FLUX_VAE_CONFIG):
    """diffusers encoder key prefix -> BFL (torchtitan ... autoencoder.Encoder) key prefix."""
    n = len(cfg["block_out_channels"])
    m = {"conv_in": "conv_in", "conv_norm_out": "norm_out", "conv_out": "conv_out",
         "mid_block.resnets.0": "mid.block_1", "mid_block.resnets.1": "mid.block_2",
         "mid_block.attentions.0.group_norm": "mid.attn_1.norm", "mid_block.attentions.0.to_q": "mid.attn_1.q",
         "mid_block.attentions.0.to_k": "mid.attn_1.k", "mid_block.attentions.0.to_v": "mid.attn_1.v",
         "mid_block.attentions.0.to_out.0": "mid.attn_1.proj_out"}
    for i in range(n):
        for j in range(cfg["layers_per_block"]):
            m[f"down_blocks.{i}.resnets.{j}"] = f"down.{i}.block.{j}"
        m[f"down_blocks.{i}.downsamplers.0.conv"] = f"down.{i}.downsample.conv"
    return m


def bfl_key_map(cfg=FLUX_VAE_CONFIG):
    """diffusers decoder key prefix -> BFL (torchtitan.experiments.flux.model.autoencoder.Decoder) key prefix."""
    n = len(cfg["block_out_channels"])
    m = {"conv_in": "conv_in", "conv_norm_out": "norm_out", "conv_out": "conv_out",
         "mid_block.resnets.0": "mid.block_1", "mid_block.resnets.1": "mid.block_2",
         "mid_block.attentions.0.group_norm": "mid.attn_1.norm", "mid_block.attentions.0.to_q": "mid.attn_1.q",
         "mid_block.attentions.0.to_k": "mid.attn_1.k", "mid_block.attentions.0.to_v": "mid.attn_1.v",
         "mid_block.attentions.0.to_out.0": "mid.attn_1.proj_out"}
    for i in range(n):
        for j in range(cfg["layers_per_block"] + 1):
            m[f"up_blocks.{i}.resnets.{j}"] = f"up.{n - 1 - i}.block.{j}"
        m[f"up_blocks.{i}.upsamplers.0.conv"] = f"up.{n - 1 - i}.upsample.conv"
    return m
