"""BASELINE config 2 shapes: FLUX-schnell 512x512 (1024 latent + 512 text tokens), 4 steps, batch B on one GPU.
python tools/bench_512.py [--batch B] [--px 512].  One JSON line: steps/s, achieved PFLOP/s, per-kernel-family CUDA-event times."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--px", type=int, default=512)
    ap.add_argument("--steps", type=int, default=16)
    args = ap.parse_args()
    import torch
    from bench import FLUX_SCHNELL, S_TXT, step_flops
    from x2i_b200 import ops
    from x2i_b200.flux import FluxTransformer2DModel
    from x2i_b200.pipeline import FluxPipeline
    dev = torch.device("cuda", 0)
    model = FluxTransformer2DModel.synthetic(FLUX_SCHNELL, device=dev, seed=0)
    B, hl = args.batch, args.px // 16
    L_img = hl * hl
    g = torch.Generator(device=dev).manual_seed(1)
    prompt = torch.randn(B, S_TXT, 4096, device=dev, generator=g).bfloat16()
    pooled = torch.randn(B, 768, device=dev, generator=g).bfloat16()
    lat = torch.randn(B, L_img, 64, device=dev, generator=g).bfloat16()
    img_ids = FluxPipeline._prepare_latent_image_ids(B, 2 * hl, 2 * hl, dev, torch.bfloat16)  # takes the LATENT height/width
    txt_ids = torch.zeros(S_TXT, 3, device=dev, dtype=torch.bfloat16)
    t = torch.full((B,), 0.75, device=dev, dtype=torch.bfloat16)

    def step():
        v = model(hidden_states=lat, timestep=t, pooled_projections=pooled, encoder_hidden_states=prompt, txt_ids=txt_ids, img_ids=img_ids,
                  return_dict=False)[0]
        ops.euler_step_(lat, v, -0.25)

    with torch.no_grad():
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    fl = step_flops(L_img, S_TXT) * B
    print(json.dumps({"workload": f"FLUX-schnell denoise step {args.px}px ({L_img}+{S_TXT} tokens), batch {B}", "ms_per_step": ms,
                      "steps_per_s": B / ms * 1e3, "tflop_per_step": fl / 1e12, "achieved_tflops": fl / ms / 1e9}))


if __name__ == "__main__":
    main()
