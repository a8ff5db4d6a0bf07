"""Sustained-load probe of the CTA-pair GEMM by epilogue kind at the shapes of the denoise step (one process, one JSON line per case):
back-to-back launches for `seconds`, SM clock sampled with nvidia-smi -> TFLOP/s, clock, tensor-pipe use implied by the clock.

    python tools/gemm_probe.py [--seconds 1.5]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.attn_probe import Sampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.5)
    a = ap.parse_args()
    from x2i_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device=dev, generator=g) * scale).bfloat16()

    M, D, H, F = 4608, 3072, 24, 12288
    x = rn(M, D)
    xf = rn(M, F)
    xc = rn(M, D + F)
    w_qkv, b_qkv = rn(3 * D, D, scale=0.02), rn(3 * D, scale=0.1)
    w_qkvm, b_qkvm = rn(3 * D + F, D, scale=0.02), rn(3 * D + F, scale=0.1)
    w_ff1, b_ff1 = rn(F, D, scale=0.02), rn(F, scale=0.1)
    w_o, b_o = rn(D, D, scale=0.02), rn(D, scale=0.1)
    w_ff2 = rn(D, F, scale=0.02)
    w_so = rn(D, D + F, scale=0.02)
    rms_q, rms_k = rn(128).abs() + 0.5, rn(128).abs() + 0.5
    ids = torch.zeros(M, 3, device=dev)
    ids[:, 1] = torch.arange(M, device=dev) // 64
    ids[:, 2] = torch.arange(M, device=dev) % 64
    rope = ops.rope_table(ids, full=False)[2]
    q, k, v = (torch.empty(1, H, M, 128, device=dev, dtype=torch.bfloat16) for _ in range(3))
    mlp = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    gate = rn(1, D)
    res = rn(M, D)
    out_qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    out_qkvm = torch.empty(M, 3 * D + F, device=dev, dtype=torch.bfloat16)
    out_f = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
    out_d = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    cases = [
        ("plain bias      N=9216  K=3072", 2.0 * M * 3 * D * D, lambda: ops.linear(x, w_qkv, b_qkv, out=out_qkv)),
        ("qkv+rms+rope    N=9216  K=3072", 2.0 * M * 3 * D * D, lambda: ops.qkv_rope(x, w_qkv, b_qkv, rms_q, rms_k, rope, q, k, v, H, M, 0)),
        ("plain bias      N=21504 K=3072", 2.0 * M * (3 * D + F) * D, lambda: ops.linear(x, w_qkvm, b_qkvm, out=out_qkvm)),
        ("qkv+rope | gelu N=21504 K=3072", 2.0 * M * (3 * D + F) * D, lambda: ops.qkv_rope(x, w_qkvm, b_qkvm, rms_q, rms_k, rope, q, k, v, H, M, 0, mlp=mlp)),
        ("gelu            N=12288 K=3072", 2.0 * M * F * D, lambda: ops.linear(x, w_ff1, b_ff1, act=1, out=out_f)),
        ("plain bias      N=3072  K=3072", 2.0 * M * D * D, lambda: ops.linear(x, w_o, b_o, out=out_d)),
        ("gate+residual   N=3072  K=3072", 2.0 * M * D * D, lambda: ops.linear_gate_residual(x, w_o, b_o, gate, res, M, out=out_d)),
        ("gate+residual   N=3072  K=12288", 2.0 * M * D * F, lambda: ops.linear_gate_residual(xf, w_ff2, b_o, gate, res, M, out=out_d)),
        ("gate+residual   N=3072  K=15360", 2.0 * M * D * (D + F), lambda: ops.linear_gate_residual(xc, w_so, b_o, gate, res, M, out=out_d)),
    ]
    for name, flops, fn in cases:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s = Sampler()
        time.sleep(0.2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        ms = e0.elapsed_time(e1) / n
        tf = flops / (ms * 1e-3) / 1e12
        peak = 148 * 8192 * (clk or 0) * 1e6 / 1e12
        print(json.dumps({"case": name, "ms": ms, "tflops_sustained": tf, "sm_mhz_median": clk, "power_w_median": pw,
                          "tensor_util_at_clock": tf / peak if peak else None}), flush=True)


if __name__ == "__main__":
    main()
