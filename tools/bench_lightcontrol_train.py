"""LightControl train step (lightcontrol/train_lightcontrol.py:672-775) at BASELINE config-5 shapes: frozen FLUX-dev,
frozen VAE, 19 trainable ControlNeXt nets, 1024x1024 style image, batch B per GPU.  One JSON line.  Under torchrun: DP over all ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--nets", type=int, default=19)
    args = ap.parse_args()
    import torch
    from bench import FLUX_SCHNELL
    from x2i_b200 import _lib, dist as xdist, train_lightcontrol as tl, vae as xv
    from x2i_b200.controlnext import ControlNeXtModel
    from x2i_b200.flux import FluxTransformer2DModel, init_synthetic_
    rank, local_rank, world = xdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = FluxTransformer2DModel.synthetic(dict(FLUX_SCHNELL, guidance_embeds=True), device=dev, seed=0).requires_grad_(False)
    if os.environ.get("X2I_CN_STREAMS"):
        model.control_net_streams = int(os.environ["X2I_CN_STREAMS"])
    vae = init_synthetic_(xv.AutoencoderKL().to(dev, torch.bfloat16).eval(), seed=5, std=0.03).requires_grad_(False)
    nets = torch.nn.ModuleList([ControlNeXtModel() for _ in range(args.nets)]).to(dev, torch.bfloat16).train()
    init_synthetic_(nets, seed=1, std=0.05)
    opt = tl.MasterWeightOptimizer(nets.parameters(), lr=1e-5, fused=True)  # fp32 masters (reference: DeepSpeed bf16 engine)
    batch = tl.synthetic_batch(args.batch, dev, seed=rank)
    torch.cuda.reset_peak_memory_stats()
    for _ in range(args.warmup):
        tl.lightcontrol_step(nets, model, vae, batch, optimizer=opt)
    torch.cuda.synchronize()
    xdist.barrier()
    n0 = _lib.launch_count()
    profiling = os.environ.get("X2I_NCU") == "1"  # ncu --profile-from-start off: capture only the timed steps
    if profiling:
        torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tl.lightcontrol_step(nets, model, vae, batch, optimizer=opt)
    e1.record()
    torch.cuda.synchronize()
    if profiling:
        torch.cuda.cudart().cudaProfilerStop()
    xdist.barrier()
    ms = xdist.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    if rank == 0:
        print(json.dumps({"workload": f"LightControl train step, FLUX-dev 1024px, {args.nets} ControlNeXt nets, VAE encode in the step",
                          "n_gpus": world, "batch_per_gpu": args.batch, "ms_per_step": ms, "samples_per_s": world * args.batch / ms * 1e3,
                          "loss": float(loss), "gpu_launches_per_step": (_lib.launch_count() - n0) / args.steps,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
