"""VAE decode at BASELINE shapes: [B, 16, 128, 128] latents -> [B, 3, 1024, 1024].
python tools/bench_vae.py [--batch B] [--steps K] [--px 1024].  One JSON line; per-kernel-family split via CUDA events."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decoder_flops(px, cfg):
    """2 * MACs of the decoder's convolutions and attention for one image."""
    boc = list(reversed(cfg["block_out_channels"]))
    h = px // 8
    f = 2 * 9 * cfg["latent_channels"] * boc[0] * h * h
    c = boc[0]
    f += 2 * (2 * 2 * 9 * c * c * h * h)                     # two mid resnets
    f += 4 * 2 * c * c * h * h + 2 * 2 * (h * h) ** 2 * c    # q,k,v,out + QK^T + PV
    prev = c
    for i, ch in enumerate(boc):
        for j in range(cfg["layers_per_block"] + 1):
            cin = prev if j == 0 else ch
            f += 2 * 9 * cin * ch * h * h + 2 * 9 * ch * ch * h * h + (2 * cin * ch * h * h if cin != ch else 0)
        prev = ch
        if i != len(boc) - 1:
            h *= 2
            f += 2 * 9 * ch * ch * h * h
    f += 2 * 9 * prev * cfg["out_channels"] * h * h
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--px", type=int, default=1024)
    args = ap.parse_args()
    import torch
    from x2i_b200 import _lib, vae as xv
    from x2i_b200.flux import init_synthetic_
    dev = torch.device("cuda", 0)
    m = xv.AutoencoderKL().to(dev, torch.bfloat16).eval()
    init_synthetic_(m, seed=5, std=0.03)
    g = torch.Generator(device=dev).manual_seed(6)
    L = (args.px // 16) ** 2
    lat = torch.randn(args.batch, L, 64, device=dev, generator=g).bfloat16()
    with torch.no_grad():
        for _ in range(args.warmup):
            xv.decode_latents(m, lat, args.px, args.px)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            img = xv.decode_latents(m, lat, args.px, args.px)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    fl = decoder_flops(args.px, xv.FLUX_VAE_CONFIG) * args.batch
    print(json.dumps({"workload": f"FLUX VAE decode {args.px}px, batch {args.batch} (infer/inference_qwenvl.py:209-216)", "ms_per_decode": ms,
                      "images_per_s": args.batch / ms * 1e3, "algorithmic_tflop_per_image": fl / args.batch / 1e12,
                      "achieved_tflops": fl / ms / 1e9, "gpu_launches_per_decode": (_lib.launch_count() - n0) / args.steps,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "finite": bool(torch.isfinite(img.float()).all())}))


if __name__ == "__main__":
    main()
