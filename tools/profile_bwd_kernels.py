"""Launches each backward kernel of the distillation step once at its FLUX shape (after one warm-up launch) so that
`ncu --set full` can capture them in isolation:  ncu --set full --clock-control none --import-source on -o ... python tools/profile_bwd_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from x2i_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
B, H, L, D, F = 1, 24, 4608, 3072, 12288
q, k, v = rn(B, H, L, 128), rn(B, H, L, 128), rn(B, H, L, 128)
do_tok = rn(B, L, D)
dy, w_out, pre = rn(L, D), rn(D, D + F) * 0.02, rn(L, F)
dbig, w_qkv = rn(L, 3 * D + F), rn(3 * D + F, D) * 0.02
x, dn, dres, sc = rn(L, D), rn(L, D), rn(L, D), rn(B, D)
stats = torch.empty(L, 2, device="cuda")
d0, d1 = torch.zeros(B, D, device="cuda"), torch.zeros(B, D, device="cuda")
wm = rn(200000, D) * 0.02
gm = torch.randn(1, 200000, device="cuda", generator=g)
dyp, xp = rn(2048, 4096), rn(2048, 4096)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    _, o1, lse = ops.attention_lse(q, k, v)
    do_hm, delta = ops.attention_bwd_prep(None, do_tok, None, o1, B, H, L, 0)
    dq, dk, dv = ops.attention_bwd(q, k, v, do_hm, lse, delta)
    ops.linear_dgrad(dy, w_out, pre=pre, n_split=D, dact=1)      # single-block proj_out dgrad + GELU'
    ops.linear_dgrad(dbig, w_qkv)                                # single-block QKV+MLP dgrad (K = 21504)
    ops.linear_wgrad(dyp, xp)                                    # projector wgrad
    ops.ln_modulate_bwd(dn, x, sc, L, dres=dres, stats=stats)
    ops.colsum(dn, B, L, out0=d0, b=x, out1=d1, stats=stats)
    ops.qk_norm_rope_bwd(dq, dk, dv, rn(L, 2 * D), rn(128), rn(128), None, dbig, L, 0)
    ops.skinny_linear_t(gm, wm)
    # KD loss forward + backward on one single-block layer pair [1, 4608, 3072], and the projector front end [1, 37, 512, 2048]
    from x2i_b200 import kd, proj as xproj
    t_kd, s_kd = rn(1, 1, L, D), rn(1, 1, L, D).requires_grad_(True)
    kd.kd_loss_stacked(t_kd, s_kd)[0].backward()
    if it == 0:
        pm = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to("cuda", torch.bfloat16)
        xin = rn(1, 37, 512, 2048)
    with torch.no_grad():
        pm(xin)
    # LightControl trainer kernels: implicit conv weight gradient (128 -> 128 and 64 -> 64 at 512 x 512, the 3x3 / stride-2 form at 256 x 256)
    # and the GroupNorm backward of a 512 x 512 x 128 layer
    if it == 0:
        xc, dyc = rn(1, 512, 512, 128), rn(1, 512, 512, 128)
        xc64, dyc64 = rn(1, 512, 512, 64), rn(1, 512, 512, 64)
        dys = rn(1, 256, 256, 128)
        gam, bet = rn(128), rn(128)
    ops.conv2d_nhwc_wgrad(xc, dyc, 3, 3, stride=1, pad=1)
    ops.conv2d_nhwc_wgrad(xc64, dyc64, 3, 3, stride=1, pad=1)
    ops.conv2d_nhwc_wgrad(xc, dys, 3, 3, stride=2, pad=1)
    ops.groupnorm_nhwc_bwd(xc, dyc, gam, bet, 4, 1e-6, act=1)
    ops.ln_modulate(x, sc, sc, L)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
