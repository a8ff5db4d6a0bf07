"""Benchmark of the X2I hot path on B200:  python bench.py --gpus N --steps K --warmup W  [--impl reference]

Metric (BASELINE.json): denoise-steps/sec at 1024x1024 (4096 latent + 512 text tokens), bf16, FLUX-schnell
architecture with random-init weights and synthetic embeddings; one "step" = one MMDiT denoise step (transformer
forward + Euler update) over the per-GPU batch.  Prints ONE JSON line (rank 0).  Keys follow the driver contract:
value (device-timed, inputs resident in HBM), e2e (through the FluxPipeline drop-in with HOST buffers: pinned H2D of the
prompt embeddings and D2H of the latents inside the timed region), roofline (fused MMDiT attention kernel, timed live
with CUDA events), cpu_baseline (oracle on the host cores, bounded sample), clocks, gpu_launches.

--impl reference times the reference's own CPU path for the same config: the PyTorch oracle restatement of the
diffusers FLUX transformer the reference calls (diffusers itself is not installable here), with all host
threads, on a bounded sample of the step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLUX_SCHNELL = dict(patch_size=1, in_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128,
                    num_attention_heads=24, joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=False,
                    axes_dims_rope=(16, 56, 56))
HEIGHT = WIDTH = 1024
S_TXT = 512
PIPE_STEPS = 4  # FLUX-schnell sampling steps (README.md:93 of the reference)
D, HEADS, FF = 3072, 24, 12288
WORKLOAD = "FLUX-schnell MMDiT denoise step, 1024x1024 (4096 latent + 512 text tokens), 19 double + 38 single blocks"


def step_flops(L_img=4096, S=S_TXT):
    L = L_img + S
    blocks = 57 * (2 * L * D * 3 * D + 2 * L * D * D + 4 * L * D * FF) + 57 * 4 * L * L * D
    emb = 2 * L_img * 64 * D * 2 + 2 * S * 4096 * D
    return blocks + emb


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(bf16=j["bf16_tflops"], bf16_sustained=j.get("bf16_tflops_sustained"), hbm=j["hbm_gbs"], source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def attn_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention kernel, from the committed ncu --set full
    capture of this same command (profiles/); None if no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_attn_traffic.json")))
    if not files:
        return None
    return json.load(open(files[-1])).get("dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
_CPU_MODEL = None


def cpu_model():
    """The full 19 + 38 block oracle transformer in bf16 on the host (24 GB).  Built on the meta device and filled by initialising
    ONE double and ONE single block and copying them into the other positions (distinct storage per block, so a step streams all
    24 GB of weights like the real model; a per-parameter random init of 11.9 B values would take minutes and times nothing)."""
    global _CPU_MODEL
    if _CPU_MODEL is not None:
        return _CPU_MODEL
    import torch
    from oracle import flux_oracle as fo
    with torch.device("meta"):
        m = fo.FluxTransformer2DModel(**FLUX_SCHNELL)
    m = m.to(torch.bfloat16).to_empty(device="cpu").eval()
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        def init(mod):
            for name, p in mod.named_parameters():
                if p.ndim >= 2:
                    p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))
                elif "norm" in name and name.endswith("weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
        for blocks in (m.transformer_blocks, m.single_transformer_blocks):
            init(blocks[0])
            src = dict(blocks[0].named_parameters())
            for b in blocks[1:]:
                for name, p in b.named_parameters():
                    p.copy_(src[name])
        for mod in (m.x_embedder, m.context_embedder, m.time_text_embed, m.norm_out, m.proj_out):
            init(mod)
    _CPU_MODEL = m
    return m


def cpu_full_step(threads):
    """ONE real denoise step of the benchmarked workload on the host cores: the oracle restatement of the diffusers FLUX transformer
    the reference calls, all 19 double + 38 single blocks at L = 512 + 4096, bf16, B = 1, plus the Euler update.  No extrapolation.
    Returns (seconds, description)."""
    import torch
    from oracle import flux_oracle as fo
    torch.set_num_threads(threads)
    m = cpu_model()
    g = torch.Generator().manual_seed(1)
    bf = torch.bfloat16
    lat = torch.randn(1, 4096, 64, generator=g).to(bf)
    prompt = torch.randn(1, S_TXT, 4096, generator=g).to(bf)
    pooled = torch.randn(1, 768, generator=g).to(bf)
    img_ids = fo.prepare_latent_image_ids(128, 128).to(bf)
    txt_ids = torch.zeros(S_TXT, 3, dtype=bf)
    with torch.no_grad():
        t0 = time.perf_counter()
        v = m(hidden_states=lat, timestep=torch.full((1,), 0.75, dtype=bf), pooled_projections=pooled, encoder_hidden_states=prompt,
              txt_ids=txt_ids, img_ids=img_ids, return_dict=False)[0]
        lat = fo.euler_step(lat, v, 0.75, 0.5)
        dt = time.perf_counter() - t0
    assert lat.shape == (1, 4096, 64)
    desc = ("oracle (PyTorch restatement of the diffusers FLUX transformer the reference calls) on CPU, bf16, B=1, L=512+4096: "
            "ONE REAL full denoise step, all 19 double + 38 single blocks + Euler update (74.4 TFLOP), no extrapolation")
    return dt, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget_s = float(os.environ.get("X2I_REF_BUDGET_S", 420))  # the whole arm stays within a few minutes of host time
    t_warm, desc = cpu_full_step(threads)                      # one warm-up step (a CPU arm measured in tens of seconds needs no more)
    n_steps = max(1, min(args.steps, int(budget_s / max(t_warm, 1e-3))))
    times = [cpu_full_step(threads)[0] for _ in range(n_steps)]
    t_step = sum(times) / len(times)
    v = 1.0 / t_step
    if n_steps < args.steps:
        desc += f"; {n_steps} of the requested {args.steps} steps timed to keep the arm within {budget_s:.0f} s"
    print(json.dumps({
        "impl": "reference", "metric": "denoise-steps/sec 1024px bf16", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 1, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": 1, "global_batch": 1, "parallelism": "cpu (rank 0 only)"},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1, help="images per GPU")
    ap.add_argument("--impl", default="x2i_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary distillation-train-step measurement")
    ap.add_argument("--train-batch", type=int, default=4, help="distillation samples per GPU (BASELINE config 4: 32 / 8 GPUs)")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the stock-PyTorch bf16 step on the same GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    from x2i_b200 import _lib, dist as xdist, ops
    from x2i_b200.flux import FluxTransformer2DModel
    from x2i_b200.pipeline import FlowMatchEulerDiscreteScheduler, FluxPipeline

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the x2i_b200 arm has no CPU fallback; use --impl reference for the CPU arm)")
    rank, local_rank, world = xdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    pk = peaks()

    # CPU baseline first (rank 0, N=1 only), before the GPU is loaded
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        global _CPU_MODEL
        cpu_full_step(threads)  # warm-up (page-in of the 24 GB host model, thread pool)
        t_cpu, desc = cpu_full_step(threads)
        cpu_base = {"value": 1.0 / t_cpu, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc}
        _CPU_MODEL = None  # release the host copy

    model = FluxTransformer2DModel.synthetic(FLUX_SCHNELL, device=dev, seed=0)
    pipe = FluxPipeline(scheduler=FlowMatchEulerDiscreteScheduler(shift=1.0), transformer=model)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    L_img = (HEIGHT // 16) * (WIDTH // 16)
    prompt = torch.randn(B, S_TXT, 4096, device=dev, generator=g).bfloat16()
    pooled = torch.randn(B, 768, device=dev, generator=g).bfloat16()
    latents = torch.randn(B, L_img, 64, device=dev, generator=g).bfloat16()
    img_ids = FluxPipeline._prepare_latent_image_ids(B, 128, 128, dev, torch.bfloat16)
    txt_ids = torch.zeros(S_TXT, 3, device=dev, dtype=torch.bfloat16)
    sched = pipe.scheduler
    sched.set_timesteps(PIPE_STEPS, device=dev)
    ts = (sched.timesteps / 1000).to(torch.bfloat16)

    def one_step(i):
        t = ts[i % PIPE_STEPS].expand(B)
        v = model(hidden_states=latents, timestep=t, pooled_projections=pooled, encoder_hidden_states=prompt, txt_ids=txt_ids,
                  img_ids=img_ids, return_dict=False)[0]
        ops.euler_step_(latents, v, -1.0 / PIPE_STEPS)

    with torch.no_grad():
        for i in range(args.warmup):
            one_step(i)
        torch.cuda.synchronize()
        xdist.barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        launches0 = _lib.launch_count()
        torch.cuda.synchronize()
        profiling = os.environ.get("X2I_NCU") == "1"  # ncu --profile-from-start off: capture only the timed region
        if profiling:
            torch.cuda.cudart().cudaProfilerStart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            one_step(i)
        e1.record()
        torch.cuda.synchronize()
        if profiling:
            torch.cuda.cudart().cudaProfilerStop()
        xdist.barrier()
        launches = _lib.launch_count() - launches0
        t_local = e0.elapsed_time(e1) * 1e-3
        t = xdist.max_over_ranks(t_local, dev)
        clk = clocks.stop() if rank == 0 else None

        # ---- e2e: the public API (FluxPipeline.__call__) with HOST buffers, per call: H2D embeds, 4 steps, D2H latents
        h_prompt = prompt.cpu().pin_memory(); h_pooled = pooled.cpu().pin_memory()
        h_lat = torch.empty(B, L_img, 64, dtype=torch.bfloat16).pin_memory()
        n_calls = max(1, args.steps // PIPE_STEPS)

        def e2e_call(seed):
            pe = h_prompt.to(dev, non_blocking=True); po = h_pooled.to(dev, non_blocking=True)
            gen = torch.Generator(device=dev).manual_seed(seed)
            out = pipe(prompt_embeds=pe, pooled_prompt_embeds=po, num_inference_steps=PIPE_STEPS, guidance_scale=3.5,
                       height=HEIGHT, width=WIDTH, output_type="latent", generator=gen).images
            h_lat.copy_(out, non_blocking=True)

        e2e_call(0)
        torch.cuda.synchronize()
        xdist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for c in range(n_calls):
            e2e_call(c + 1)
        s1.record()
        torch.cuda.synchronize()
        xdist.barrier()
        t_e2e = xdist.max_over_ranks(s0.elapsed_time(s1) * 1e-3, dev)
        h2d = (h_prompt.numel() + h_pooled.numel()) * 2 / PIPE_STEPS
        d2h = h_lat.numel() * 2 / PIPE_STEPS

        # ---- roofline of the named kernel (fused MMDiT attention): every one of its launches inside real denoise steps
        # is bracketed with CUDA events on the launch stream (eager mode, same kernels/buffers as the timed region, L2 state
        # as in the pipeline: q,k,v were just written by the QKV GEMM); average over 2 steps x 57 launches.
        roof = None
        if rank == 0:
            L = S_TXT + L_img
            model.use_cuda_graph = False
            one_step(0)
            ops.ATTN_EVENTS = []
            one_step(1)
            one_step(2)
            torch.cuda.synchronize()
            durs = [a.elapsed_time(b) * 1e-3 for a, b in ops.ATTN_EVENTS]
            ops.ATTN_EVENTS = None
            model.use_cuda_graph = True
            t_att = sum(durs) / len(durs)
            fl = 4.0 * L * L * 128 * HEADS * B
            ach = fl / t_att / 1e12
            pk_s = pk.get("bf16_sustained") or pk["bf16"]
            # the same kernel and the library kernel it replaces (torch SDPA -> cuDNN / flash on this box), each timed ALONE on the
            # same q, k, v (back-to-back launches, L2-warm, burst clocks): explains the in-step number, not a bench value
            import torch.nn.functional as F_
            ws = model._ws
            q_, k_, v_ = ws["q"], ws["k"], ws["v"]
            o0 = torch.empty(B, S_TXT, D, device=dev, dtype=torch.bfloat16)
            o1 = torch.empty(B, L_img, D, device=dev, dtype=torch.bfloat16)

            def timed(fn, n=50, settle_s=0.5):
                """Same protocol for both kernels: launch back to back for `settle_s` first, so the power cap and the SM clock
                are in the state the kernel itself produces (a burst after an idle gap runs ~25 % faster and measures the host's
                pause, not the kernel), then time n launches."""
                fn()
                torch.cuda.synchronize()
                t_end = time.perf_counter() + settle_s
                while time.perf_counter() < t_end:
                    for _ in range(20):
                        fn()
                    torch.cuda.synchronize()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b_.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b_) * 1e-3 / n

            t_iso = timed(lambda: ops.attention(q_, k_, v_, split=S_TXT, out0=o0, out1=o1))
            t_lib = timed(lambda: F_.scaled_dot_product_attention(q_, k_, v_))
            roof = {"kernel": ("mmdit_attention_fwd_persistent_kernel<2, 0, %d>" % int(os.environ.get("X2I_ATTN_LAG", "1") != "0")) if os.environ.get("X2I_ATTN_PERSIST", "1") == "1" else "mmdit_attention_fwd_kernel<2, 0, 0>", "bound": "tensor", "achieved": ach, "peak": pk_s,
                    "unit": "TFLOP/s", "frac": ach / pk_s, "frac_of_burst_peak": ach / pk["bf16"],
                    "peak_source": pk["source"] + " (sustained cuBLAS bf16: the kernel is timed inside long denoise steps)",
                    "ms_per_launch": t_att * 1e3, "launches_timed": len(durs), "algorithmic_flops_per_launch": fl, "traffic": attn_traffic(),
                    "step_share_attention": 57 * t_att / (t_local / args.steps),
                    "isolated_tflops": fl / t_iso / 1e12, "isolated_frac_of_sustained_peak": fl / t_iso / 1e12 / pk_s,
                    "library_tflops": fl / t_lib / 1e12,
                    "library": "torch.nn.functional.scaled_dot_product_attention (torch 2.11 backend choice on sm_100), same q/k/v; both "
                               "kernels timed alone under their own sustained (power-capped) load: 0.5 s of back-to-back launches, then 50 timed"}

        # ---- GPU library baseline: the oracle restatement of the reference's model in bf16 with stock PyTorch ops on
        # this same GPU (F.scaled_dot_product_attention + cuBLAS + eager elementwise): the number X2I's own code path reaches here.
        lib_base = None
        if rank == 0 and world == 1 and not args.no_library_baseline:
            from oracle import flux_oracle as fo
            with torch.device("meta"):
                om = fo.FluxTransformer2DModel(**FLUX_SCHNELL)
            om = om.to(torch.bfloat16).to_empty(device=dev).eval()
            om.load_state_dict(model.state_dict())
            lat2 = latents.clone()

            def lib_step(i):
                nonlocal lat2
                v = om(hidden_states=lat2, timestep=ts[i % PIPE_STEPS].expand(B), pooled_projections=pooled, encoder_hidden_states=prompt,
                       txt_ids=txt_ids, img_ids=img_ids, return_dict=False)[0]
                lat2 = fo.euler_step(lat2, v, 0.0, -1.0 / PIPE_STEPS)

            for i in range(3):
                lib_step(i)
            torch.cuda.synchronize()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            n_lib = max(4, min(args.steps, 8))
            for i in range(n_lib):
                lib_step(i)
            l1.record()
            torch.cuda.synchronize()
            t_libstep = l0.elapsed_time(l1) * 1e-3 / n_lib
            lib_base = {"value": B / t_libstep, "unit": "steps/s", "ms_per_step": t_libstep * 1e3, "steps": n_lib,
                        "kind": "oracle restatement of the diffusers FLUX transformer, bf16, stock PyTorch eager ops on the same B200 "
                                "(F.scaled_dot_product_attention + cuBLAS + elementwise kernels), same weights and inputs, device-timed",
                        "speedup_of_x2i_b200": (t_libstep) / (t_local / args.steps)}
            del om, lat2
            torch.cuda.empty_cache()

    # ---- secondary workload (BASELINE config 4, N=1 only): the attention-distillation train step on the same frozen FLUX --
    # teacher pass + projector + student pass (saving mode) + KD loss + backward through all 57 blocks + projector wgrad +
    # AdamW.  Reported as an extra object; the headline metric above is unchanged.  tools/bench_train.py runs it under torchrun.
    train_info = None
    if not args.no_train:
        from x2i_b200 import proj as xproj, train as xtrain
        model.use_cuda_graph = False
        TB = args.train_batch
        proj = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to(dev, torch.bfloat16)
        opt = torch.optim.AdamW(proj.parameters(), lr=1e-4, fused=True)
        bucket = xdist.GradBucket(proj.parameters())  # all projector gradients in one flat buffer: ONE all-reduce, no copies
        tb = xtrain.synthetic_batch(TB, dev, FLUX_SCHNELL, seed=7 + rank)
        for _ in range(2):
            xtrain.distill_step(proj, model, tb, optimizer=opt, bucket=bucket)
        torch.cuda.synchronize()
        xdist.barrier()
        n0 = _lib.launch_count()
        timings = []
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        n_train = 3
        for _ in range(n_train):
            tl = xtrain.distill_step(proj, model, tb, optimizer=opt, bucket=bucket, timings=timings)
        t1e.record()
        torch.cuda.synchronize()
        xdist.barrier()
        tt = xdist.max_over_ranks(t0e.elapsed_time(t1e) * 1e-3, dev) / n_train
        ar_ms = xdist.max_over_ranks(sum(a.elapsed_time(b) for a, b in timings) / max(1, len(timings)), dev) if timings else 0.0
        train_info = {"workload": "attention-distillation train step (train_qwenvl.py:559-651 + teacher :717-816), 1024px latents, "
                                  "projector qwen3b [B,37,512,2048], teacher+student on every GPU, DP over all ranks "
                                  "(BASELINE config 4: 4 samples per GPU)",
                      "batch_per_gpu": TB, "global_batch": TB * world, "n_gpus": world, "ms_per_step": tt * 1e3,
                      "samples_per_s": TB * world / tt, "scaling": "weak",
                      "approx_tflops_per_gpu": TB * step_flops() * 4.3 / tt / 1e12, "loss": float(tl),
                      "gpu_launches_per_step": (_lib.launch_count() - n0) / n_train,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "collective": {"op": "NCCL all_reduce(AVG) of the projector gradients, one per optimizer step (train_qwenvl.py:483)",
                                     "bytes": bucket.nbytes, "ms": ar_ms,
                                     "share_of_step": ar_ms / (tt * 1e3) if tt > 0 else None}}
        model.use_cuda_graph = True
        del proj, opt, tb, bucket
        torch.cuda.empty_cache()

    # ---- secondary workload (SURVEY 8f N2, N=1 only): the VAE decode that follows the 4 denoise steps of an image
    vae_info = None
    if world == 1 and not args.no_train:
        from x2i_b200 import vae as xvae
        from x2i_b200.flux import init_synthetic_
        vae = init_synthetic_(xvae.AutoencoderKL().to(dev, torch.bfloat16).eval(), seed=5, std=0.03)
        with torch.no_grad():
            for _ in range(2):
                xvae.decode_latents(vae, latents, HEIGHT, WIDTH)
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(3):
                xvae.decode_latents(vae, latents, HEIGHT, WIDTH)
            v1.record()
            torch.cuda.synchronize()
        vae_info = {"workload": "FLUX VAE decode of the step's latents -> [B,3,1024,1024] (infer/inference_qwenvl.py:209-216)",
                    "ms_per_decode": v0.elapsed_time(v1) / 3, "algorithmic_tflop_per_image": 10.47}
        del vae
        torch.cuda.empty_cache()

    # ---- secondary workloads (SURVEY 8f N3 / N4 and BASELINE config 5, N=1 only): the step right before the path and the LightControl branch.
    # Extra objects of the same line so the driver's record carries them; each is a few seconds.  --no-train skips them too.
    mllm_info = lc_info = None
    if world == 1 and not args.no_train:
        from x2i_b200 import mllm as xmllm, proj as xproj
        del model
        torch.cuda.empty_cache()
        with torch.no_grad():
            mm = xmllm.Qwen2_5_VLTextPrefill.synthetic(xmllm.QWEN2_5_VL_3B, device=dev, seed=0)
            pm = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to(dev, torch.bfloat16)
            gen = torch.Generator(device=dev).manual_seed(1)
            ids = torch.randint(0, 151000, (1, S_TXT), device=dev, generator=gen)
            mask = torch.ones(1, S_TXT, dtype=torch.long, device=dev)
            mask[:, :300] = 0  # a 212-token prompt, left-padded to 512 like the reference's processor call
            te = torch.empty(1, 37, S_TXT, 2048, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                pm(mm.prefill_hidden_states(ids, mask, out=te))
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for _ in range(10):
                pm(mm.prefill_hidden_states(ids, mask, out=te))
            m1.record()
            torch.cuda.synchronize()
        mllm_info = {"workload": "Qwen2.5-VL-3B text prefill with all-layer capture [1,37,512,2048] + Proj7Exp (infer/inference_qwenvl.py:176-179)",
                     "ms_per_prompt": m0.elapsed_time(m1) / 10}
        del mm, pm, te
        torch.cuda.empty_cache()
        from x2i_b200 import train_lightcontrol as tl, vae as xvae2
        from x2i_b200.controlnext import ControlNeXtModel
        from x2i_b200.flux import init_synthetic_ as init_syn
        dev_model = FluxTransformer2DModel.synthetic(dict(FLUX_SCHNELL, guidance_embeds=True), device=dev, seed=0).requires_grad_(False)
        vae2 = init_syn(xvae2.AutoencoderKL().to(dev, torch.bfloat16).eval(), seed=5, std=0.03).requires_grad_(False)
        nets = torch.nn.ModuleList([ControlNeXtModel() for _ in range(19)]).to(dev, torch.bfloat16).train()
        init_syn(nets, seed=1, std=0.05)
        opt2 = tl.MasterWeightOptimizer(nets.parameters(), lr=1e-5, fused=True)
        lb = tl.synthetic_batch(1, dev, seed=0)
        tl.lightcontrol_step(nets, dev_model, vae2, lb, optimizer=opt2)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(2):
            ll = tl.lightcontrol_step(nets, dev_model, vae2, lb, optimizer=opt2)
        l1.record()
        torch.cuda.synchronize()
        lc_info = {"workload": "LightControl train step (lightcontrol/train_lightcontrol.py:672-775): FLUX-dev 1024px, 19 trainable ControlNeXt nets, "
                               "VAE encode in the step, B = 1",
                   "ms_per_step": l0.elapsed_time(l1) / 2, "loss": float(ll)}
        del dev_model, vae2, nets, opt2, lb
        torch.cuda.empty_cache()

    if rank == 0:
        total_steps = world * B * args.steps
        value = total_steps / t
        e2e_v = world * B * n_calls * PIPE_STEPS / t_e2e
        fl = step_flops()
        out = {
            "metric": "denoise-steps/sec 1024px bf16", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init FLUX-schnell weights, N(0,1) embeddings/latents)",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world} (replicas, no collective)",
                       "l2_policy": "per-step working set (24 GB weights) >> 126 MB L2; no flush needed"},
            "tflops_per_gpu": fl * B * args.steps / t_local / 1e12,
            "step_roofline_frac_bf16": fl * B * args.steps / t_local / 1e12 / pk["bf16_sustained"] if pk.get("bf16_sustained") else None,
            "e2e": {"value": e2e_v, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "x2i_b200.pipeline.FluxPipeline.__call__ (4-step schnell sampling per call, pinned host buffers)"},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu_base, "gpu_library_baseline": lib_base, "clocks": clk,
            "distill_train": train_info, "vae_decode": vae_info, "mllm_prefill": mllm_info, "lightcontrol_train": lc_info,
        }
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
