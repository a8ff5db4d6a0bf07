"""LightControl editing step (BASELINE config 5 shapes): FLUX-dev denoise step at 1024px with 19 ControlNeXt nets on a
1024x1024 hint.  python tools/bench_lightcontrol.py [--batch B] [--steps K].  One JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    from bench import FLUX_SCHNELL, step_flops
    from x2i_b200 import _lib, ops
    from x2i_b200.controlnext import ControlNeXtModel
    from x2i_b200.flux import FluxTransformer2DModel, init_synthetic_
    from x2i_b200.pipeline import FluxPipeline
    dev = torch.device("cuda", 0)
    cfg = dict(FLUX_SCHNELL, guidance_embeds=True)
    model = FluxTransformer2DModel.synthetic(cfg, device=dev, seed=0)
    nets = torch.nn.ModuleList([ControlNeXtModel() for _ in range(19)]).to(dev, torch.bfloat16).eval()
    init_synthetic_(nets, seed=1, std=0.05)
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(2)
    r = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()  # noqa: E731
    prompt, pooled, lat = r(B, 512, 4096), r(B, 768), r(B, 4096, 64)
    hint = (torch.rand(B, 3, 1024, 1024, device=dev, generator=g) * 2 - 1).bfloat16()
    img_ids = FluxPipeline._prepare_latent_image_ids(B, 128, 128, dev, torch.bfloat16)
    txt_ids = torch.zeros(512, 3, device=dev, dtype=torch.bfloat16)
    t = torch.full((B,), 0.7, device=dev, dtype=torch.bfloat16)
    gd = torch.full((B,), 3.5, device=dev, dtype=torch.bfloat16)

    def step(with_control):
        v = model(hidden_states=lat, timestep=t, guidance=gd, pooled_projections=pooled, encoder_hidden_states=prompt, txt_ids=txt_ids,
                  img_ids=img_ids, guided_hint=hint if with_control else None, control_nets=nets if with_control else None,
                  return_dict=False)[0]
        ops.euler_step_(lat, v, -1.0 / 20)

    res = {}
    with torch.no_grad():
        for tag, wc in (("with_control", True), ("plain", False)):
            for _ in range(args.warmup):
                step(wc)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step(wc)
            e1.record()
            torch.cuda.synchronize()
            res[tag] = (e0.elapsed_time(e1) / args.steps, (_lib.launch_count() - n0) / args.steps)
        # the 19 nets alone
        x = torch.zeros(B, 4096, 3072, device=dev, dtype=torch.bfloat16)
        for _ in range(2):
            for n in nets:
                n.forward_tokens(hint, t * 1000, add_to=x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for n in nets:
            n.forward_tokens(hint, t * 1000, add_to=x)
        e1.record()
        torch.cuda.synchronize()
        t_nets = e0.elapsed_time(e1)
        # the same 19 nets through ControlNeXtStack (one launch per layer) + the per-net last conv with the fused injection
        from x2i_b200.controlnext import ControlNeXtStack
        stack = ControlNeXtStack(nets)
        for _ in range(2):
            mids = stack.mid_features(hint, t * 1000)
            for i, n in enumerate(nets):
                n.finish_tokens(mids[i], add_to=x)
        torch.cuda.synchronize()
        e0.record()
        mids = stack.mid_features(hint, t * 1000)
        for i, n in enumerate(nets):
            n.finish_tokens(mids[i], add_to=x)
        e1.record()
        torch.cuda.synchronize()
        t_stack = e0.elapsed_time(e1)
    conv_flops = 19 * 436.8e9 * B
    print(json.dumps({"workload": "LightControl editing denoise step: FLUX-dev 1024px + 19 ControlNeXt nets on a 1024x1024 hint",
                      "batch": B, "ms_per_step_with_control": res["with_control"][0], "ms_per_step_plain": res["plain"][0],
                      "steps_per_s_with_control": B * 1e3 / res["with_control"][0], "launches_per_step": res["with_control"][1],
                      "controlnext_19_nets_per_net_ms": t_nets, "controlnext_19_nets_stacked_ms": t_stack,
                      "controlnext_stacked_tflops": conv_flops / (t_stack * 1e-3) / 1e12,
                      "step_tflops_with_control": (step_flops() * B + conv_flops) / (res["with_control"][0] * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
