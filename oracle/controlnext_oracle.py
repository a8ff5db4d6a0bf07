"""ORACLE (test infrastructure only) for the LightControl editing branch: ``ControlNeXtModel`` and its injection into
the FLUX double blocks.

Restates ``/root/reference/lightcontrol/lightcontrol_flux.py``:
  * ``ControlNeXtModel`` ctor :580-668 and forward :708-749 (in-tree; PINNED: oracle/make_golden.py imports the reference
    class itself with the two diffusers leaves below injected and stores its outputs in tests/golden/controlnext.pt);
  * the injection ``hidden_states += control_nets[i](guided_hint, timestep)['out'].flatten(2).transpose(1, 2) * scale`` after
    each of the first ``len(control_nets)`` double blocks (:504-507).
The leaves ``ResnetBlock2D`` and ``Downsample2D`` live in ``diffusers==0.31.0`` (``models/resnet.py``, ``models/downsampling.py``;
un-vendored, requirements.txt:3) and are restated from the published algorithm for exactly the configuration
the reference instantiates: ``ResnetBlock2D(in_channels, out_channels, temb_channels, groups)`` with the library defaults
(eps 1e-6, SiLU, time_embedding_norm "default", output_scale_factor 1, 1x1 conv shortcut when the channel count changes)
and ``Downsample2D(channels, use_conv=True, out_channels, padding=1, name="op")`` (3x3 stride-2 conv stored as ``conv``).
PARITY OF THESE TWO LEAVES IS UNPINNED by the reference itself.
State-dict keys follow diffusers so a trained ``controlnet.state_dict()`` (train_lightcontrol.py:785-791) loads.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .flux_oracle import TimestepEmbedding, Timesteps


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, temb_channels=512, groups=32, eps=1e-6, **_unused):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0) if in_channels != out_channels else None

    def forward(self, x, temb, *args, **kwargs):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv", **_unused):
        super().__init__()
        assert use_conv, "the reference only instantiates the conv form"
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x, *args, **kwargs):
        return self.conv(x)


class ControlNeXtModel(nn.Module):
    """lightcontrol_flux.py:575-749."""

    def __init__(self, in_channels=(128, 128), out_channels=(128, 256), groups=(4, 8), time_embed_dim=256, final_out_channels=320,
                 hidden_out=3072):
        super().__init__()
        self.time_proj = Timesteps(128, True, downscale_freq_shift=0)
        self.time_embedding = TimestepEmbedding(128, time_embed_dim)
        self.embedding = nn.Sequential(
            nn.Conv2d(3, 64, 3, stride=2, padding=1), nn.GroupNorm(2, 64), nn.ReLU(),
            nn.Conv2d(64, 64, 3, padding=1), nn.GroupNorm(2, 64), nn.ReLU(),
            nn.Conv2d(64, 128, 3, padding=1), nn.GroupNorm(2, 128), nn.ReLU())
        self.down_res = nn.ModuleList([ResnetBlock2D(in_channels=i, out_channels=o, temb_channels=time_embed_dim, groups=g)
                                       for i, o, g in zip(in_channels, out_channels, groups)])
        self.down_sample = nn.ModuleList([Downsample2D(o, use_conv=True, out_channels=o, padding=1, name="op") for o in out_channels])
        c = out_channels[-1]
        self.mid_convs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, c, 3, padding=1), nn.ReLU(), nn.GroupNorm(8, c), nn.Conv2d(c, c, 3, padding=1), nn.GroupNorm(8, c)),
            nn.Conv2d(c, hidden_out, kernel_size=2, stride=2)])
        self.scale = 1.0

    def forward(self, sample, timestep):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(sample.dtype))
        sample = self.embedding(sample)
        for res, down in zip(self.down_res, self.down_sample):
            sample = down(res(sample, emb), emb)
        sample = self.mid_convs[0](sample) + sample
        return {"out": self.mid_convs[1](sample), "scale": self.scale}


def inject(hidden_states, control):
    """lightcontrol_flux.py:505-507."""
    out = control["out"].flatten(2).transpose(1, 2).to(hidden_states.dtype)
    return hidden_states + out * control["scale"]
