"""ORACLE (test infrastructure only) for the X2I alignment projector.

Restates ``/root/reference/utils/proj.py``: ``MLP3`` :14-33 and ``Proj7Exp`` :35-72 with the
factories :74-96 (``use_t5=False`` only -- the T5 branch of the reference raises NameError,
SURVEY.md Appendix C.1).  PINNED: oracle/make_golden.py imports the reference module itself,
copies its weights into this restatement and stores reference outputs in
tests/golden/proj_*.pt; tests/test_oracle_golden.py checks this file against them.
State-dict keys are identical to the reference's (conv.weight, conv.bias | cha_scale,
mlp.layernorm.*, mlp.projector.{0,2}.weight, mlp.fc.1.{weight,bias}).
"""
import torch
import torch.nn as nn


class MLP3(nn.Module):
    def __init__(self, in_dim=4096, out_dim=4096, hidden_dim=4096, out_dim1=768, layer_norm_eps=1e-5):
        super().__init__()
        self.layernorm = nn.LayerNorm(in_dim, eps=layer_norm_eps)
        self.projector = nn.Sequential(nn.Linear(in_dim, hidden_dim, bias=False), nn.GELU(),
                                       nn.Linear(hidden_dim, hidden_dim, bias=False))
        self.fc = nn.Sequential(nn.GELU(), nn.Linear(out_dim, out_dim1))

    def forward(self, x):
        seq = self.projector(self.layernorm(x))
        pooled = self.fc(seq).mean(dim=1)
        return pooled, seq


class Proj7Exp(nn.Module):
    def __init__(self, in_channels=25, kernel_size=5, input_dim=896, output_dim0=768, output_dim1=4096,
                 norm_eps=1e-6, use_scale=True, use_cnn=True):
        super().__init__()
        self.use_scale, self.use_cnn = use_scale, use_cnn
        if use_scale:
            self.cha_scale = nn.Parameter(torch.empty(1, in_channels, 1, 1))
            nn.init.xavier_normal_(self.cha_scale, gain=1)
        elif use_cnn:
            self.conv = nn.Conv2d(in_channels, 1, kernel_size=kernel_size, padding=(kernel_size - 1) // 2)
        self.mlp = MLP3(input_dim, output_dim1, output_dim1, output_dim0, norm_eps)

    def forward(self, x):
        B, C, S, H = x.shape
        if self.use_scale:
            x = (self.cha_scale * x).mean(dim=1)
        elif self.use_cnn:
            x = self.conv(x).squeeze(1)
        else:
            x = x.mean(dim=1)
        return self.mlp(x)


_DIMS = {"qwen3b": 2048, "qwen7b": 3584, "internvl1b": 896, "internvl4b": 2048, "minicpm": 3584}


def create_proj(kind: str, in_channels: int, use_scale: bool, use_cnn: bool) -> Proj7Exp:
    """kind in qwen3b|qwen7b|internvl1b|internvl4b|minicpm (utils/proj.py:74-96).  The qwen*/minicpm
    factories force use_cnn=False when use_scale is set (:75,:80,:94)."""
    if kind in ("qwen3b", "qwen7b", "minicpm") and use_scale:
        use_cnn = False
    return Proj7Exp(in_channels=in_channels, kernel_size=5, input_dim=_DIMS[kind], output_dim0=768,
                    output_dim1=4096, norm_eps=1e-6, use_scale=use_scale, use_cnn=use_cnn)
