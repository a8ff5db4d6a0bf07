"""Probe of the tensor-core layer-mixing convolution (x2i_proj_mix_ln_tc) against the FP32-pipe stencil kernel and an fp32 torch reference.
Env: X2I_PROJCONV_ALIGNED=0|1 (60- or 56-column tiles), X2I_PROJCONV_BASE_OFFSET=1|0 (descriptor base-offset field).  One JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from x2i_b200 import ops  # noqa: E402


def main():
    res = {}
    for (B, C, S, H) in ((1, 37, 512, 2048), (2, 29, 128, 3584), (1, 3, 128, 512), (4, 37, 512, 2048)):
        g = torch.Generator(device="cuda").manual_seed(B + C)
        x = torch.randn(B, C, S, H, device="cuda", generator=g).bfloat16()
        w = (torch.randn(C, 25, device="cuda", generator=g) * 0.1).bfloat16().float()
        gamma = (1 + 0.1 * torch.randn(H, device="cuda", generator=g)).float()
        beta = (0.1 * torch.randn(H, device="cuda", generator=g)).float()
        ref = torch.nn.functional.conv2d(x.float(), w.view(1, C, 5, 5), torch.tensor([0.3], device="cuda"), padding=2)[:, 0]
        refn = torch.nn.functional.layer_norm(ref, (H,), gamma, beta, 1e-6)
        ops.proj_conv_tensor_cores = True
        y, xm = ops.proj_mix_ln_save(x, 0, w, 0.3, gamma, beta, 1e-6)
        ops.proj_conv_tensor_cores = False
        y0, xm0 = ops.proj_mix_ln_save(x, 0, w, 0.3, gamma, beta, 1e-6)
        rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
        tag = f"{B}x{C}x{S}x{H}"
        res[tag] = dict(tc_xm=rel(xm, ref), tc_y=rel(y, refn), stencil_xm=rel(xm0, ref), stencil_y=rel(y0, refn), tc_vs_stencil=rel(y, y0))
        for flag, name in ((True, "tc_ms"), (False, "stencil_ms")):
            ops.proj_conv_tensor_cores = flag
            for _ in range(3):
                ops.proj_mix_ln(x, 0, w, 0.3, gamma, beta, 1e-6)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.proj_mix_ln(x, 0, w, 0.3, gamma, beta, 1e-6)
            e1.record()
            torch.cuda.synchronize()
            res[tag][name] = e0.elapsed_time(e1) / 20
    print(json.dumps(res))


if __name__ == "__main__":
    main()
