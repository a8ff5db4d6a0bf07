"""Sustained-load probe of the fused attention forward (one process per variant: the env switches are read once per process).

    X2I_ATTN_PAIR=0|1 python tools/attn_probe.py [--sdpa] [--seconds 2.0] [--bwd]

Launches the kernel back to back for `seconds` while nvidia-smi samples SM clock and power, then prints one JSON line:
TFLOP/s, median SM clock under load, power, and the tensor-pipe utilisation implied by the clock
(4096 bf16 FMA / clk / SM x 148 SMs).  Separates "latency-bound at a high clock" from "power-bound at a low clock"."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Sampler:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._r, daemon=True).start()

    def _r(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        self.p.terminate()
        clk, pw = [], []
        for t, ln in self.lines:
            if t0 + 0.3 <= t <= t1:
                try:
                    a, b = ln.split(",")
                    clk.append(float(a)); pw.append(float(b))
                except ValueError:
                    pass
        clk.sort(); pw.sort()
        return (clk[len(clk) // 2] if clk else None), (pw[len(pw) // 2] if pw else None), len(clk)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sdpa", action="store_true")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--L", type=int, default=4608)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    from x2i_b200 import ops
    B, H, L = a.B, 24, a.L
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, H, L, 128, device="cuda", generator=g).bfloat16() for _ in range(3))
    split = 512 if L > 512 else 0
    o0 = torch.empty(B, split, H * 128, device="cuda", dtype=torch.bfloat16) if split else None
    o1 = torch.empty(B, L - split, H * 128, device="cuda", dtype=torch.bfloat16)
    flops = 4.0 * L * L * 128 * H * B
    if a.bwd:
        flops *= 2.5
        if a.sdpa:
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
            do = torch.randn_like(q)

            def fn():
                o = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
                o.backward(do)
            flops = flops / 2.5 * 3.5  # forward + backward
        else:
            a0, a1, lse = ops.attention_lse(q, k, v, split=split)
            do = torch.randn(B, L, H * 128, device="cuda", generator=g).bfloat16()
            do_hm, delta = ops.attention_bwd_prep(do[:, :split].contiguous() if split else None, do[:, split:].contiguous(), a0, a1, B, H, L, split)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)

            def fn():
                ops.attention_bwd(q, k, v, do_hm, lse, delta, dq=dq, dk=dk, dv=dv)
    elif a.sdpa:
        def fn():
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
    else:
        def fn():
            ops.attention(q, k, v, split=split, out0=o0, out1=o1)
    # parity spot check (forward variants)
    err = None
    if not a.sdpa and not a.bwd:
        fn()
        o = torch.cat([t for t in (o0, o1) if t is not None], 1).view(B, L, H, 128).transpose(1, 2)
        errs = []
        for h in (0, 11, 23):
            ref = torch.nn.functional.scaled_dot_product_attention(q[:, h:h + 1].float(), k[:, h:h + 1].float(), v[:, h:h + 1].float())
            errs.append(float((o[:, h:h + 1].float() - ref).norm() / ref.norm()))
        err = max(errs)
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    # burst: 20 launches from a cool state are not available here (the GPU is warm); report the first 20 anyway
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    burst = flops * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    s = Sampler()
    time.sleep(0.2)
    t0 = time.time()
    n = 0
    e0.record()
    while time.time() - t0 < a.seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    clk, pw, ns = s.stop(t0, t1)
    tf = flops * n / (e0.elapsed_time(e1) * 1e-3) / 1e12
    peak_at_clk = 4096 * 2 * 148 * clk * 1e6 / 1e12 if clk else None
    print(json.dumps({"variant": a.tag or ("sdpa" if a.sdpa else f"x2i pair={os.environ.get('X2I_ATTN_PAIR', '1')}"), "bwd": a.bwd, "B": B, "L": L,
                      "tflops_sustained": tf, "tflops_first20": burst, "ms_per_launch": flops / tf / 1e9, "sm_mhz_median": clk, "power_w_median": pw,
                      "clock_samples": ns, "tensor_peak_at_clock": peak_at_clk, "tensor_util_at_clock": tf / peak_at_clk if peak_at_clk else None,
                      "rel_err_vs_fp32_sdpa": err}))


if __name__ == "__main__":
    main()
