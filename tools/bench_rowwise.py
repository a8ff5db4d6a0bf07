"""HBM-bound row-wise kernels at their FLUX / projector shapes: achieved GB/s of algorithmic bytes against the measured copy bandwidth
(MEASURED_PEAKS.json).  Each kernel is timed over inputs larger than L2 (rotating buffers) with CUDA events.  One JSON line per kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def timed(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / n


def main():
    from x2i_b200 import kd, ops, proj as xproj
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
    L, D = 4608, 3072
    out = []
    # KD loss: one layer pair per launch (what kd_loss_layers issues, 76 per step) and a 19-layer stacked group
    NB = 8  # rotating buffers: 8 x 2 x 28 MB > L2
    ts, ss = [rn(1, 1, L, D) for _ in range(NB)], [rn(1, 1, L, D).requires_grad_(True) for _ in range(NB)]
    t_f = timed(lambda i: kd.kd_loss_stacked(ts[i % NB], ss[i % NB].detach()), 40)
    out.append(dict(kernel="kd_row_kernel<fwd> + reductions, 1 layer [1,4608,3072]", ms=t_f * 1e3, gbs=2 * L * D * 2 / t_f / 1e9))

    def fb(i):
        s_ = ss[i % NB]
        s_.grad = None
        kd.kd_loss_stacked(ts[i % NB], s_)[0].backward()
    t_fb = timed(fb, 40)
    out.append(dict(kernel="kd loss fwd + bwd, 1 layer [1,4608,3072]", ms=t_fb * 1e3, gbs=5 * L * D * 2 / t_fb / 1e9))
    tb, sb = rn(1, 19, L, D), rn(1, 19, L, D).requires_grad_(True)
    t_f19 = timed(lambda i: kd.kd_loss_stacked(tb, sb.detach()), 10)
    out.append(dict(kernel="kd_row_kernel<fwd> + reductions, 19 layers stacked [1,19,4608,3072]", ms=t_f19 * 1e3, gbs=2 * 19 * L * D * 2 / t_f19 / 1e9))

    def fb19(i):
        sb.grad = None
        kd.kd_loss_stacked(tb, sb)[0].backward()
    t_fb19 = timed(fb19, 10)
    out.append(dict(kernel="kd loss fwd + bwd, 19 layers stacked", ms=t_fb19 * 1e3, gbs=5 * 19 * L * D * 2 / t_fb19 / 1e9))
    # LN + modulate at the single-block shape
    xs, ys = [rn(L, D) for _ in range(NB)], [torch.empty(L, D, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    sc, sh = rn(1, D), rn(1, D)
    t_ln = timed(lambda i: ops.ln_modulate(xs[i % NB], sc, sh, L, out=ys[i % NB]), 100)
    out.append(dict(kernel="ln_modulate_kernel [4608,3072]", ms=t_ln * 1e3, gbs=2 * L * D * 2 / t_ln / 1e9))
    # projector front end: 5x5 layer-mixing conv + LayerNorm, [1,37,512,2048]
    pm = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to("cuda", torch.bfloat16)
    xin = [rn(1, 37, 512, 2048) for _ in range(3)]
    w = pm.conv.weight.detach().float().reshape(37, 25).contiguous()
    m = pm.mlp
    ga, be, cb = m.layernorm.weight.detach().float(), m.layernorm.bias.detach().float(), float(pm.conv.bias.detach().float())
    t_pm = timed(lambda i: ops.proj_mix_ln(xin[i % 3], 0, w, cb, ga, be, m.layernorm.eps), 20)
    out.append(dict(kernel="proj_conv_tc_kernel + ln_rows_f32_kernel [1,37,512,2048] (5x5 conv over layers on the tensor pipe + LN)", ms=t_pm * 1e3, gbs=37 * 512 * 2048 * 2 / t_pm / 1e9,
                    gflops=2 * 37 * 25 * 512 * 2048 / t_pm / 1e9))
    ops.proj_conv_tensor_cores = False
    t_st = timed(lambda i: ops.proj_mix_ln(xin[i % 3], 0, w, cb, ga, be, m.layernorm.eps), 20)
    ops.proj_conv_tensor_cores = True
    out.append(dict(kernel="proj_mix_ln_kernel [1,37,512,2048] (the FP32-pipe stencil it replaces: round 1)", ms=t_st * 1e3, gbs=37 * 512 * 2048 * 2 / t_st / 1e9,
                    gflops=2 * 37 * 25 * 512 * 2048 / t_st / 1e9))
    for o in out:
        o["frac_of_measured_hbm"] = o["gbs"] / pk
        print(json.dumps(o))


if __name__ == "__main__":
    main()
