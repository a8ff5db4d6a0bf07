"""Small-shape launches of the hand-rolled mbarrier / TMEM pipelines for compute-sanitizer:

    compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize_small.py

attention forward (self + ragged tail + cross form), attention backward (both launches), the CTA-pair grouped GEMM with the QKV /
gate-residual epilogues, the single-CTA GEMM, the implicit-GEMM conv.  Results are checked loosely (the tool is the checker)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from x2i_b200 import ops  # noqa: E402


def rn(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*s, device="cuda", generator=g) * scale).bfloat16()


def main():
    # attention forward + lse, ragged L (not a multiple of 128 / 256), split outputs
    B, H, L = 1, 2, 300
    q, k, v = rn(B, H, L, 128, seed=1), rn(B, H, L, 128, seed=2), rn(B, H, L, 128, seed=3)
    o0, o1, lse = ops.attention_lse(q, k, v, split=44)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    o = torch.cat([o0, o1], 1).view(B, L, H, 128).permute(0, 2, 1, 3)
    print("attn fwd rel", float((o.float() - ref).norm() / ref.norm()))
    # lagged soft-max steps with the overflow redo pass (scores growing by ~2^167 per key tile): drain, barrier re-initialisation, second pass
    Lr = 640
    ramp = (torch.arange(Lr, device="cuda", dtype=torch.float32) * 0.02).view(1, 1, Lr, 1)
    qr = torch.full((1, 2, Lr, 128), 4.0, device="cuda").bfloat16()
    kr = ramp.expand(1, 2, Lr, 128).bfloat16().contiguous()
    vr = rn(1, 2, Lr, 128, seed=21)
    orr = torch.empty(1, Lr, 256, device="cuda", dtype=torch.bfloat16)
    ops.attention(qr, kr, vr, split=0, out1=orr)
    refr = torch.nn.functional.scaled_dot_product_attention(qr.float(), kr.float(), vr.float())
    print("attn fwd (redo pass) rel", float((orr.view(1, Lr, 2, 128).permute(0, 2, 1, 3).float() - refr).norm() / refr.norm()))
    # backward
    do = rn(B, L, H * 128, seed=4)
    do_hm, delta = ops.attention_bwd_prep(do[:, :44].contiguous(), do[:, 44:].contiguous(), o0, o1, B, H, L, 44)
    dq, dk, dv = ops.attention_bwd(q, k, v, do_hm, lse, delta)
    print("attn bwd finite", bool(torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()))
    # the 2-CTA cluster form of the backward (TMA multicast of the streamed tiles, multicast commits; odd tile count -> padding tile)
    os.environ["X2I_ATTN_BWD_MC"] = "1"
    dq2, dk2, dv2 = ops.attention_bwd(q, k, v, do_hm, lse, delta)
    del os.environ["X2I_ATTN_BWD_MC"]
    print("attn bwd cluster form bit-identical", bool(torch.equal(dq, dq2) and torch.equal(dk, dk2) and torch.equal(dv, dv2)))
    # cross-attention form with key-padding lengths (resampler)
    qc, kc, vc = rn(2, 2, 64, 128, seed=5), rn(2, 2, 200, 128, seed=6), rn(2, 2, 200, 128, seed=7)
    oc = ops.cross_attention(qc, kc, vc, kv_len=torch.tensor([200, 77], device="cuda", dtype=torch.int32))
    print("cross attn finite", bool(torch.isfinite(oc.float()).all()))
    # GEMMs: CTA-pair kernel (M > 128, N % 256 == 0) and single-CTA kernel, bias + GELU epilogue
    x, w, b = rn(384, 256, seed=8), rn(512, 256, seed=9, scale=0.05), rn(512, seed=10, scale=0.1)
    y = ops.linear(x, w, b, act=1)
    yr = torch.nn.functional.gelu(x.float() @ w.float().T + b.float(), approximate="tanh")
    print("gemm2 rel", float((y.float() - yr).norm() / yr.norm()))
    x1, w1 = rn(96, 128, seed=11), rn(192, 128, seed=12, scale=0.05)
    y1 = ops.linear(x1, w1, None)
    print("gemm rel", float((y1.float() - x1.float() @ w1.float().T).norm() / (x1.float() @ w1.float().T).norm()))
    # gate + residual epilogue with the un-gated copy (hook tensor)
    gate, res = rn(2, 512, seed=13), rn(384, 512, seed=14)
    aux = torch.empty(384, 512, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(res)
    ops.linear_gate_residual(x, w, b, gate, res, 192, out=out, aux=aux)
    print("gate-residual finite", bool(torch.isfinite(out.float()).all()))
    # dgrad / wgrad operand forms
    dy = rn(384, 512, seed=15)
    dx = ops.linear_dgrad(dy, w)
    dw = ops.linear_wgrad(dy, x)
    print("dgrad/wgrad finite", bool(torch.isfinite(dx.float()).all() and torch.isfinite(dw.float()).all()))
    # implicit-GEMM conv
    xi, wi, bi = rn(1, 16, 24, 64, seed=16), rn(128, 64, 3, 3, seed=17, scale=0.05), rn(128, seed=18, scale=0.1)
    yc = ops.conv2d_nhwc(xi, ops.pack_conv_weight(wi), bi, 3, 3, stride=1, pad=1)
    print("conv finite", bool(torch.isfinite(yc.float()).all()))
    # implicit conv weight gradient (tap-shifted boxes as the wgrad GEMM's B operand), stride 1 and the stride-2 parity view
    xw, dyw = rn(2, 8, 64, 64, seed=19), rn(2, 8, 64, 128, seed=20)
    dw1, _ = ops.conv2d_nhwc_wgrad(xw, dyw, 3, 3, stride=1, pad=1)
    xs, dys = rn(1, 16, 128, 64, seed=21), rn(1, 8, 64, 64, seed=22)
    dw2, _ = ops.conv2d_nhwc_wgrad(xs, dys, 3, 3, stride=2, pad=1)
    print("implicit conv wgrad finite", bool(torch.isfinite(dw1.float()).all() and torch.isfinite(dw2.float()).all()))
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
