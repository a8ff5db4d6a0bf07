"""CPU: host-side mirror of the reference interfaces (no kernels are launched here)."""
import os

import numpy as np
import pytest
import torch

from oracle import flux_oracle as fo
from x2i_b200 import dist as xdist
from x2i_b200._lib import X2IError
from x2i_b200.flux import FluxTransformer2DModel
from x2i_b200.pipeline import FlowMatchEulerDiscreteScheduler, FluxPipeline, calculate_shift
from x2i_b200 import proj as xproj


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_pipeline_static_helpers_bit_exact(golden_dir):
    g = _load(golden_dir, "helpers.pt")
    assert torch.equal(FluxPipeline._prepare_latent_image_ids(2, 8, 12, "cpu", torch.float32), g["ids_8x12"])
    assert torch.equal(FluxPipeline._prepare_latent_image_ids(1, 128, 128, "cpu", torch.float32), g["ids_128"])
    assert torch.equal(FluxPipeline._pack_latents(g["pack_in"], 2, 16, 8, 12), g["pack_out"])
    assert torch.equal(FluxPipeline._unpack_latents(g["pack_out"], 8 * 8, 12 * 8, 16), g["pack_in"])
    assert calculate_shift(4096) == g["shift_4096"] and calculate_shift(1024) == g["shift_1024"]
    assert calculate_shift(4096, 256, 4096, 0.5, 1.15) == g["shift_dev_4096"]


@pytest.mark.parametrize("dyn", [False, True])
def test_scheduler_matches_oracle(dyn):
    s = FlowMatchEulerDiscreteScheduler(shift=1.0, use_dynamic_shifting=dyn)
    mu = calculate_shift(4096, 256, 4096, 0.5, 1.15)
    for n in (1, 4, 20, 28):
        s.set_timesteps(sigmas=np.linspace(1.0, 1 / n, n), mu=mu if dyn else None)
        ref = fo.flow_match_sigmas(n, mu if dyn else None, 1.0, dyn)
        assert torch.equal(s.sigmas, ref)
        assert torch.equal(s.timesteps, ref[:-1] * 1000)
    s.set_timesteps(sigmas=[1.0], mu=mu if dyn else None)  # distillation: one step at t = 1000 (train_qwenvl.py:753-771)
    assert float(s.timesteps[0]) == 1000.0


def test_state_dict_keys_are_diffusers_names():
    cfg = dict(num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=1, joint_attention_dim=16,
               pooled_projection_dim=8, in_channels=8, guidance_embeds=True)
    with torch.device("meta"):
        mine = FluxTransformer2DModel(**cfg)
        ref = fo.FluxTransformer2DModel(**cfg)
    a, b = mine.state_dict(), ref.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape, k
    for k in ("transformer_blocks.0.attn.to_out.0.weight", "transformer_blocks.0.ff.net.0.proj.weight",
              "single_transformer_blocks.0.proj_out.weight", "time_text_embed.guidance_embedder.linear_1.weight",
              "transformer_blocks.0.attn.norm_added_k.weight", "norm_out.linear.bias"):
        assert k in a


def test_attn_processor_plugin_surface():
    with torch.device("meta"):
        m = FluxTransformer2DModel(num_layers=2, num_single_layers=3, num_attention_heads=1, joint_attention_dim=16,
                                   pooled_projection_dim=8, in_channels=8)
    procs = m.attn_processors
    assert len(procs) == 5 and "transformer_blocks.0.attn.processor" in procs
    sentinel = object()
    m.set_attn_processor(sentinel)
    assert all(p is sentinel for p in m.attn_processors.values())
    with pytest.raises(ValueError):
        m.set_attn_processor({"x": sentinel})
    m.set_attn_processor({k: i for i, k in enumerate(procs)})
    assert m.single_transformer_blocks[2].attn.get_processor() == 4
    # hooks attach to a real nn.Module (train_qwenvl.py:206-214)
    h = m.transformer_blocks[0].attn.register_forward_hook(lambda mod, i, o: None)
    h.remove()


def test_no_cpu_fallback():
    m = FluxTransformer2DModel(num_layers=1, num_single_layers=1, num_attention_heads=1, joint_attention_dim=16,
                               pooled_projection_dim=8, in_channels=8)
    with torch.no_grad(), pytest.raises(X2IError):
        m(hidden_states=torch.randn(1, 4, 8), encoder_hidden_states=torch.randn(1, 2, 16), pooled_projections=torch.randn(1, 8),
          timestep=torch.ones(1), img_ids=torch.zeros(4, 3), txt_ids=torch.zeros(2, 3))
    p = xproj.create_proj_internvl4b(5, use_t5=False, use_scale=False, use_cnn=True)
    with torch.no_grad(), pytest.raises(X2IError):
        p(torch.randn(1, 5, 4, 2048))


def test_projector_factories_match_reference_shapes(golden_dir):
    g = _load(golden_dir, "proj_c1.pt")
    with torch.device("meta"):
        p = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True)
        assert sum(x.numel() for x in p.parameters()) == g["n_params"]
        assert set(p.state_dict()) == {"conv.weight", "conv.bias", "mlp.layernorm.weight", "mlp.layernorm.bias",
                                       "mlp.projector.0.weight", "mlp.projector.2.weight", "mlp.fc.1.weight", "mlp.fc.1.bias"}
        q = xproj.create_proj3_qwen7b(29, use_t5=False, use_scale=True, use_cnn=True)
        assert "cha_scale" in q.state_dict() and not q.use_cnn  # use_scale wins (utils/proj.py:80)
        assert xproj.create_proj_internvl1b(25, use_t5=False).mlp.layernorm.normalized_shape == (896,)
    with pytest.raises(NameError):
        xproj.create_proj_minicpm(29)  # use_t5=True default: dead branch of the reference (NameError there too)


def test_projector_checkpoint_prefix_stripping():
    p = xproj.Proj7Exp(in_channels=3, input_dim=16, output_dim0=8, output_dim1=16, use_t5=False, use_scale=False, use_cnn=True)
    sd = {"module." + k: torch.randn_like(v) for k, v in p.state_dict().items()}
    xproj.load_projector_state(p, sd)
    assert torch.equal(p.conv.weight, sd["module.conv.weight"])


def test_shard_range_covers_batch():
    for n in (1, 7, 8, 16, 33):
        for w in (1, 2, 4, 8):
            spans = [xdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_resampler_surface_and_pos_table(golden_dir):
    from x2i_b200.resampler import Resampler, get_2d_sincos_pos_embed
    g = _load(golden_dir, "resampler_small.pt")
    m = Resampler(num_queries=8, embed_dim=256, num_heads=2, kv_dim=48, adaptive=True, max_size=(6, 7))
    assert set(m.state_dict()) == set(g["state"])          # same checkpoint keys as the reference module
    m.load_state_dict(g["state"])
    assert torch.equal(m.pos_embed, g["pos_embed"])          # bit-exact sincos table
    assert "pos_embed" not in m.state_dict()                 # non-persistent buffer, like the reference
    assert get_2d_sincos_pos_embed(8, (2, 3)).shape == (2, 3, 8)
    with pytest.raises(X2IError):
        Resampler(num_queries=8, embed_dim=96, num_heads=2)   # head_dim must be 128
    with pytest.raises(X2IError):
        m(torch.randn(3, 14, 48), g["tgt_sizes"])              # no CPU path


def test_projector_bundle_and_checkpoint_formats(tmp_path):
    """On-disk contracts: the trainer's {step}/diffusion_pytorch_model.bin with DDP 'module.' prefixes
    (train_qwenvl.py:641-647, inference_qwenvl.py:85-91) and the ComfyUI {"config","state_dict"} bundle (x2i_comfyui/model.py:33-39)."""
    import torch
    from x2i_b200 import proj as xproj, train
    p = xproj.create_proj_internvl1b(5, use_t5=False, use_scale=False, use_cnn=True)
    fn = train.save_projector_checkpoint(p, str(tmp_path), 1300)
    assert fn.endswith("1300/diffusion_pytorch_model.bin")
    sd = torch.load(fn)
    q = xproj.create_proj_internvl1b(5, use_t5=False, use_scale=False, use_cnn=True)
    xproj.load_projector_state(q, {"module." + k: v for k, v in sd.items()})
    for a, b in zip(p.state_dict().values(), q.state_dict().values()):
        assert torch.equal(a, b)
    bundle = str(tmp_path / "proj.pt")
    xproj.save_projector_bundle(p, bundle)
    r = xproj.load_projector_bundle(bundle, device=None)
    assert type(r) is xproj.Proj7Exp and r.use_cnn and not r.use_scale
    for a, b in zip(p.state_dict().values(), r.state_dict().values()):
        assert torch.equal(a, b)
    d = torch.load(bundle, weights_only=True)
    assert set(d) == {"config", "state_dict"} and d["config"]["in_channels"] == 5 and d["config"]["input_dim"] == 896


def test_bench_reference_arm_json_contract(monkeypatch, capsys):
    """bench.py --impl reference prints ONE JSON line with the driver's keys (the CPU sample itself is mocked: it takes minutes)."""
    import argparse
    import json
    import bench
    monkeypatch.setattr(bench, "cpu_full_step", lambda threads: (12.5, "mocked full step"))
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(argparse.Namespace(gpus=1, steps=2, warmup=1))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "steps/s" and abs(j["value"] - 1 / 12.5) < 1e-9 and j["higher_is_better"] is True
    assert j["steps"] == 2  # every requested step is a real, fully timed step while the arm fits its time budget
    assert j["config"]["workload"] == bench.WORKLOAD and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    monkeypatch.setenv("RANK", "1")  # under torchrun only rank 0 prints
    bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=1))
    assert capsys.readouterr().out.strip() == ""


def test_control_net_stack_eligibility():
    """ControlNeXtStack only takes over for >= 2 x2i_b200 ControlNeXtModels of identical layout; anything else keeps the per-net
    protocol of lightcontrol_flux.py:504-507."""
    import torch.nn as nn
    from x2i_b200.controlnext import ControlNeXtModel, ControlNeXtStack
    a, b = ControlNeXtModel(), ControlNeXtModel()
    assert ControlNeXtStack.supported([a, b]) and ControlNeXtStack.supported(nn.ModuleList([a, b]))
    assert not ControlNeXtStack.supported([a])
    assert not ControlNeXtStack.supported([a, nn.Identity()])
    assert not ControlNeXtStack.supported([a, ControlNeXtModel(out_channels=(128, 128))])

    class Sub(ControlNeXtModel):  # a subclass may override forward: not stacked
        pass

    assert not ControlNeXtStack.supported([a, Sub()])


def test_lightcontrol_host_helpers():
    """Host-side pieces of the LightControl train step (lightcontrol/train_lightcontrol.py:690-703): logit-normal timestep density,
    the training sigma table (FLUX.1-dev scheduler config: dynamic shifting on -> unshifted; static shift 3 when off), and the
    no-CPU-path rule."""
    import pytest
    import torch
    from x2i_b200 import train_lightcontrol as tl
    from x2i_b200._lib import X2IError
    raw = torch.linspace(1.0, 1.0 / 1000, 1000)
    assert torch.equal(tl.train_sigmas(1000, 3.0), raw)                          # default: use_dynamic_shifting -> no static shift
    s = tl.train_sigmas(1000, 3.0, use_dynamic_shifting=False)
    assert s.shape == (1000,) and abs(float(s[0]) - 1.0) < 1e-6 and bool((s[1:] < s[:-1]).all())
    assert torch.allclose(s, 3.0 * raw / (1 + 2.0 * raw))
    assert torch.allclose(tl.train_sigmas(1000, 1.0, use_dynamic_shifting=False), raw)   # shift 1 = the schnell table
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    u1 = tl.compute_density_for_timestep_sampling("logit_normal", 64, 0.0, 1.0, generator=g1)
    u2 = tl.compute_density_for_timestep_sampling("logit_normal", 64, 0.0, 1.0, generator=g2)
    assert torch.equal(u1, u2) and float(u1.min()) > 0 and float(u1.max()) < 1
    g3 = torch.Generator().manual_seed(5)
    assert torch.allclose(u1, torch.sigmoid(torch.randn(64, generator=g3)))
    um = tl.compute_density_for_timestep_sampling("mode", 64, generator=torch.Generator().manual_seed(1))
    assert um.shape == (64,)
    with pytest.raises(X2IError):
        tl.lightcontrol_step(None, None, None, {"pixel_values": torch.zeros(1, 3, 16, 16)})
    b = tl.synthetic_batch(1, "cpu", height=32, width=48, S=4, seed=0)
    assert b["pixel_values"].shape == (1, 3, 32, 48) and b["prompt_embeds"].shape == (1, 4, 4096) and float(b["pixel_values"].abs().max()) <= 1


def test_master_weight_optimizer_keeps_small_updates():
    """at the reference's lr = 1e-5 an Adam update is below half a bf16 ulp of a 0.02-0.05 weight; stepping bf16 parameters
    directly loses it, fp32 masters (DeepSpeed bf16 engine, accelerate_config_debug.yaml) keep it."""
    from x2i_b200.train_lightcontrol import MasterWeightOptimizer
    torch.manual_seed(0)
    w0 = (torch.randn(64, 64) * 0.04).bfloat16()
    direct = torch.nn.Parameter(w0.clone())
    shadow = torch.nn.Parameter(w0.clone())
    opt_d = torch.optim.AdamW([direct], lr=1e-5, weight_decay=0.0)
    opt_m = MasterWeightOptimizer([shadow], lr=1e-5, weight_decay=0.0)
    g = torch.ones_like(w0)
    for _ in range(40):  # constant gradient: Adam moves every weight by ~lr per step -> 4e-4 in total (several bf16 ulps of 0.04)
        direct.grad = g.clone(); opt_d.step(); opt_d.zero_grad()
        shadow.grad = g.clone(); opt_m.step(); opt_m.zero_grad()
    moved_direct = float((direct.detach().float() - w0.float()).abs().mean())
    moved_master = float((opt_m.master[0].detach() - w0.float()).abs().mean())
    assert abs(moved_master - 40e-5) < 2e-5                      # the masters followed Adam exactly
    assert moved_direct < 0.2 * moved_master                     # pure bf16 stalls (most updates round away)
    assert abs(float((shadow.detach().float() - w0.float()).mean()) + 40e-5) < 1.5e-4  # bf16 weights track the masters to within rounding
    sd = opt_m.state_dict()
    opt_m.load_state_dict(sd)
    assert torch.equal(shadow.detach(), opt_m.master[0].bfloat16())


def test_grad_bucket_views_allreduce_and_clip():
    torch.manual_seed(1)
    lin = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    bucket = xdist.GradBucket(lin.parameters())
    assert bucket.flat.numel() == sum(p.numel() for p in lin.parameters()) and bucket.nbytes == 4 * bucket.flat.numel()
    x = torch.randn(5, 8)
    lin(x).pow(2).sum().backward()
    lin(x).pow(2).sum().backward()          # accumulation lands in the same buffer (train_qwenvl.py:561: grads add up over micro-steps)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4))
    ref.load_state_dict(lin.state_dict())
    (2 * ref(x).pow(2).sum()).backward()
    flat_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
    assert torch.allclose(bucket.flat, flat_ref, rtol=1e-5, atol=1e-6)
    bucket.allreduce_mean_()                # world size 1: no-op
    total = bucket.clip_grad_norm_(1.0)
    want = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    assert abs(float(total) - float(want)) < 1e-4 * float(want)
    assert torch.allclose(bucket.flat, torch.cat([p.grad.reshape(-1) for p in ref.parameters()]), rtol=1e-4, atol=1e-7)
    bucket.zero_()
    assert float(bucket.flat.abs().max()) == 0 and lin[0].weight.grad.data_ptr() == bucket.flat.data_ptr()


def test_auto_resume_picks_highest_numbered_checkpoint(tmp_path):
    """train_qwenvl.py:199-203,:404-410,:534-536: weights-only resume from {output_dir}/{max step}/diffusion_pytorch_model.bin."""
    from x2i_b200 import train
    assert train.get_max_numbered_filename(str(tmp_path)) is None and train.resume_projector(None, str(tmp_path / "missing")) is None
    p = xproj.Proj7Exp(in_channels=3, input_dim=64, use_t5=False, use_scale=False, use_cnn=True)
    for step, val in ((500, 0.25), (1500, 0.5), (1000, 0.75)):
        with torch.no_grad():
            p.conv.bias.fill_(val)
        train.save_projector_checkpoint(p, str(tmp_path), step)
    q = xproj.Proj7Exp(in_channels=3, input_dim=64, use_t5=False, use_scale=False, use_cnn=True)
    assert train.resume_projector(q, str(tmp_path)) == 1500
    assert float(q.conv.bias) == 0.5


def test_lightcontrol_return_form_serves_both_call_sites():
    """diffusers returns (sample,) for return_dict=False (train_qwenvl.py:587 indexes [0]); the vendored LightControl
    transformer returns the bare tensor and train_lightcontrol.py:745-751 hands it straight to _unpack_latents."""
    from x2i_b200.flux import _TensorTuple
    t = torch.arange(2 * 6 * 64, dtype=torch.float32).view(2, 6, 64)
    r = _TensorTuple((t,))
    assert isinstance(r, tuple) and len(r) == 1 and r[0] is t
    assert r.shape == t.shape and r.dtype == t.dtype
    assert torch.equal(FluxPipeline._unpack_latents(r, 32, 48, 16), FluxPipeline._unpack_latents(t, 32, 48, 16))
