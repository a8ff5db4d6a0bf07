"""Distillation train step (BASELINE config 4 shapes) on one or N GPUs: python tools/bench_train.py [--batch B] [--steps K]
Prints one JSON line: samples/s, ms/step, phase breakdown, peak memory.  Under torchrun: DP over all ranks (one all-reduce)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLUX_DEV = dict(patch_size=1, in_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128,
                num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True,
                axes_dims_rope=(16, 56, 56))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--layers", type=int, nargs=2, default=None, help="override (double, single) block counts")
    args = ap.parse_args()
    import torch
    from x2i_b200 import _lib, dist as xdist, proj as xproj, train
    from x2i_b200.flux import FluxTransformer2DModel
    rank, local_rank, world = xdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = dict(FLUX_DEV)
    if args.layers:
        cfg["num_layers"], cfg["num_single_layers"] = args.layers
    model = FluxTransformer2DModel.synthetic(cfg, device=dev, seed=0).requires_grad_(False)
    proj = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to(dev, torch.bfloat16)
    opt = torch.optim.AdamW(proj.parameters(), lr=1e-4, fused=True)
    batch = train.synthetic_batch(args.batch, dev, cfg, seed=rank)
    torch.cuda.reset_peak_memory_stats()
    for _ in range(args.warmup):
        train.distill_step(proj, model, batch, optimizer=opt)
    torch.cuda.synchronize()
    xdist.barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = train.distill_step(proj, model, batch, optimizer=opt)
    e1.record()
    torch.cuda.synchronize()
    xdist.barrier()
    t = xdist.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev) / args.steps
    # phase breakdown (rank 0, one extra step with events between phases)
    phases = {}
    if rank == 0:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        S = batch["prompt_embeds_t5"].shape[1]
        dt = model.dtype
        kw = dict(hidden_states=batch["latents"], timestep=batch["timestep"] / 1000, txt_ids=torch.zeros(S, 3, device=dev, dtype=dt),
                  img_ids=train.prepare_latent_image_ids(128, 128, dev, dt), guidance=torch.full((args.batch,), 3.5, device=dev, dtype=dt))
        ev[0].record()
        kt, ks = [], []
        with torch.no_grad():
            train._run_hooked(model, kt, encoder_hidden_states=batch["prompt_embeds_t5"], pooled_projections=batch["pooled_clip"], **kw)
        ev[1].record()
        a, e = proj(batch["text_embeddings"])
        train._run_hooked(model, ks, encoder_hidden_states=e, pooled_projections=a, **kw)
        from x2i_b200 import kd
        l, _ = kd.kd_loss_layers(kt[0] + kt[1] + kt[2], ks[0] + ks[1] + ks[2])
        ev[2].record()
        l.backward()
        ev[3].record()
        torch.cuda.synchronize()
        phases = {"teacher_fwd_ms": ev[0].elapsed_time(ev[1]), "student_fwd_loss_ms": ev[1].elapsed_time(ev[2]),
                  "backward_ms": ev[2].elapsed_time(ev[3])}
        proj.zero_grad()
    if rank == 0:
        flops = 74.38e12 * (cfg["num_layers"] + cfg["num_single_layers"]) / 57
        print(json.dumps({"workload": "attention-distillation train step, FLUX-dev 1024px (4096+512 tokens), projector qwen3b C=37 S=512 H=2048",
                          "n_gpus": world, "batch_per_gpu": args.batch, "blocks": [cfg["num_layers"], cfg["num_single_layers"]],
                          "ms_per_step": t * 1e3, "samples_per_s": world * args.batch / t, "loss": float(loss),
                          "approx_tflops_per_gpu": args.batch * flops * (1 + 1 + 2.3) / t / 1e12, **phases,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "gpu_launches_per_step": (_lib.launch_count() - n0) / args.steps}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
