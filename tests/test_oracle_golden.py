"""CPU: the oracle restatements against the fixtures minted FROM THE REFERENCE by oracle/make_golden.py."""
import json
import os

import pytest
import torch

from oracle import flux_oracle as fo
from oracle import kd_oracle, proj_oracle


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_projector_small_all_branches(golden_dir):
    g = _load(golden_dir, "proj_small.pt")
    for tag, kw in (("cnn", dict(use_scale=False, use_cnn=True)), ("scale", dict(use_scale=True, use_cnn=False)),
                    ("mean", dict(use_scale=False, use_cnn=False))):
        m = proj_oracle.Proj7Exp(in_channels=5, kernel_size=5, input_dim=64, output_dim0=24, output_dim1=96, **kw)
        m.load_state_dict(g[tag]["state"])  # identical keys to the reference module
        with torch.no_grad():
            pooled, seq = m(g[tag]["x"])
        torch.testing.assert_close(pooled, g[tag]["pooled"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(seq, g[tag]["seq"], rtol=1e-5, atol=1e-6)


def test_projector_config1(golden_dir):
    """BASELINE.json configs[0]: projector fwd + MSE vs random T5 embeds, batch=1 seq=77, CPU."""
    from oracle.make_golden import synth_state
    g = _load(golden_dir, "proj_c1.pt")
    m = proj_oracle.create_proj("qwen3b", 37, use_scale=False, use_cnn=True)
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] == 28317342
    m.load_state_dict(synth_state(m, 21, std=0.02))
    x = torch.randn(1, 37, 77, 2048, generator=torch.Generator().manual_seed(0))
    t5 = torch.randn(1, 77, 4096, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        pooled, seq = m(x)
    torch.testing.assert_close(pooled, g["pooled"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(seq[:, ::7, ::13], g["seq_sub"], rtol=1e-4, atol=1e-5)
    assert abs(float(torch.nn.functional.mse_loss(seq, t5)) - g["mse"]) < 1e-5


def test_train_helpers(golden_dir):
    g = _load(golden_dir, "helpers.pt")
    torch.testing.assert_close(kd_oracle.normalize(g["norm_in"]), g["norm_out"], rtol=0, atol=0)
    assert torch.equal(fo.prepare_latent_image_ids(8, 12), g["ids_8x12"])
    assert torch.equal(fo.prepare_latent_image_ids(128, 128), g["ids_128"])
    assert torch.equal(fo.pack_latents(g["pack_in"]), g["pack_out"])
    assert fo.calculate_shift(4096) == g["shift_4096"]
    assert fo.calculate_shift(1024) == g["shift_1024"]
    assert fo.calculate_shift(4096, 256, 4096, 0.5, 1.15) == g["shift_dev_4096"]
    # unpack is the inverse of pack (train_lightcontrol.py:403-410)
    lat = g["pack_in"]
    assert torch.equal(fo.unpack_latents(fo.pack_latents(lat), 8 * 8, 12 * 8, 16), lat)


def test_flux_structure_matches_reference_classes(golden_dir):
    """Oracle transformer == the reference's in-tree block/transformer classes (same leaves)."""
    g = _load(golden_dir, "flux_structure.pt")
    for tag in ("schnell", "dev"):
        e = g[tag]
        m = fo.FluxTransformer2DModel(**e["cfg"]).eval()
        m.load_state_dict(e["state"])  # diffusers key names
        hooks = [[], [], []]

        def two(mod, i, o, hooks=hooks):
            hooks[0].append(o[0]); hooks[1].append(o[1])

        def one(mod, i, o, hooks=hooks):
            hooks[2].append(o)

        for b in m.transformer_blocks:
            b.attn.register_forward_hook(two)
        for b in m.single_transformer_blocks:
            b.attn.register_forward_hook(one)
        with torch.no_grad():
            y = m(**e["inputs"], return_dict=False)[0]
        torch.testing.assert_close(y, e["output"], rtol=1e-5, atol=1e-6)
        for mine, ref in zip(hooks, e["hooks"]):
            torch.testing.assert_close(torch.stack(mine, 1), ref, rtol=1e-5, atol=1e-6)


def test_torchtitan_crosscheck_recorded(golden_dir):
    r = json.load(open(os.path.join(golden_dir, "crosscheck.json")))
    if r.get("available"):
        assert max(r["double_img_maxabs"], r["double_txt_maxabs"], r["single_maxabs"]) < 1e-5
        assert r["timestep_sinusoid_maxabs"] == 0.0


def test_kd_loss_direction_and_guard():
    g = torch.Generator().manual_seed(3)
    t = [torch.randn(2, 5, 48, generator=g) for _ in range(3)]
    s = [torch.randn(2, 5, 48, generator=g) for _ in range(3)]
    loss, skipped = kd_oracle.kd_loss(t, s)
    assert skipped == [] and float(loss) > 0
    # KL(student || teacher): zero iff identical
    assert abs(float(kd_oracle.kd_loss(t, t)[0])) < 1e-6
    # explicit formula
    def term(a, b):
        pa = torch.softmax(kd_oracle.normalize(a) / 3, -1)
        pb = torch.softmax(kd_oracle.normalize(b) / 3, -1)
        return (pb * (pb.log() - pa.log())).sum() / a.shape[0]
    torch.testing.assert_close(loss, sum(term(a, b) for a, b in zip(t, s)), rtol=1e-5, atol=1e-6)
    # a layer producing nan is skipped, not propagated (train_qwenvl.py:606-609)
    s[1][0, 0, 0] = float("nan")
    loss2, skipped2 = kd_oracle.kd_loss(t, s)
    assert skipped2 == [1] and torch.isfinite(loss2)


def test_rope_identity_on_text_tokens():
    cos, sin = fo.rope_table(torch.zeros(4, 3))
    assert torch.equal(cos, torch.ones(4, 128)) and torch.equal(sin, torch.zeros(4, 128))
    ids = fo.prepare_latent_image_ids(8, 8)
    cos, sin = fo.rope_table(ids)
    assert torch.equal(cos[:, :16], torch.ones(16, 16))  # axis 0 is always 0


def test_resampler_oracle_matches_reference(golden_dir):
    from oracle import resampler_oracle as ro
    g = _load(golden_dir, "resampler_small.pt")
    m = ro.Resampler(num_queries=8, embed_dim=256, num_heads=2, kv_dim=48, adaptive=True, max_size=(6, 7)).eval()
    m.load_state_dict(g["state"])
    assert torch.equal(m.pos_embed, g["pos_embed"])  # index-derived table: bit-exact
    with torch.no_grad():
        y = m(g["x"], g["tgt_sizes"])
    torch.testing.assert_close(y, g["out"], rtol=1e-5, atol=1e-6)


def test_adaln_leaves_match_an_independent_in_image_copy():
    """Second anchor for the un-vendored diffusers AdaLN leaves (parity is unpinned by the reference itself): the
    F5-TTS-derived DiT inside transformers' Qwen2.5-Omni carries the same two modules (chunk order shift/scale/gate x2, and the
    scale-FIRST order of the final AdaLayerNormContinuous).  Same weights -> identical outputs."""
    omni = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    if not hasattr(omni, "Qwen2_5_OmniAdaLayerNormZero"):
        pytest.skip("this transformers build has no Qwen2_5_OmniAdaLayerNormZero")
    from oracle import flux_oracle as fo
    torch.manual_seed(0)
    dim = 64
    mine, theirs = fo.AdaLayerNormZero(dim).eval(), omni.Qwen2_5_OmniAdaLayerNormZero(dim).eval()
    theirs.load_state_dict({k: v for k, v in mine.state_dict().items() if k.startswith("linear.")}, strict=False)
    x, emb = torch.randn(2, 7, dim), torch.randn(2, dim)
    with torch.no_grad():
        a, b = mine(x, emb), theirs(x, emb)
    assert len(a) == len(b) == 5
    for u, v in zip(a, b):
        assert torch.allclose(u, v, atol=1e-6)
    fin, fin_t = fo.AdaLayerNormContinuous(dim, dim).eval(), omni.Qwen2_5_OmniAdaLayerNormZero_Final(dim).eval()
    fin_t.load_state_dict({k: v for k, v in fin.state_dict().items() if k.startswith("linear.")}, strict=False)
    with torch.no_grad():
        assert torch.allclose(fin(x, emb), fin_t(x, emb), atol=1e-6)


def test_lightcontrol_trainer_helpers_match_reference(golden_dir):
    """The index / layout helpers of lightcontrol/train_lightcontrol.py (:383-:422), pinned by a fixture minted from the reference
    file itself (oracle/make_golden.py::golden_lightcontrol_helpers): bit-exact pack / unpack / ids, and the trainer's sigma table
    equals what the reference's get_sigmas() looks up in the scheduler."""
    from x2i_b200 import train_lightcontrol as tl
    from x2i_b200.pipeline import FluxPipeline
    d = torch.load(os.path.join(golden_dir, "lightcontrol_helpers.pt"))
    assert torch.equal(FluxPipeline._pack_latents(d["lat"], 2, 16, 8, 12), d["packed"])
    assert torch.equal(FluxPipeline._unpack_latents(d["packed"], 64, 96, 16), d["unpacked"])
    assert torch.equal(d["unpacked"], d["lat"])                                        # round trip
    assert torch.equal(FluxPipeline._prepare_latent_image_ids(2, 8, 12, "cpu", torch.float32), d["ids"])
    assert torch.equal(fo.unpack_latents(d["packed"], 64, 96, 16), d["unpacked"]) and torch.equal(fo.pack_latents(d["lat"]), d["packed"])
    # the reference's scheduler comes from FLUX.1-dev's scheduler_config.json (use_dynamic_shifting = true -> NO static shift)
    table = tl.train_sigmas_from_config(d["scheduler_config"])
    assert torch.equal(table, tl.train_sigmas()) and float(table[0]) == 1.0 and abs(float(table[500]) - 0.5) < 1e-6
    assert torch.allclose(table[d["idx"]], d["sigmas"].flatten(), rtol=0, atol=1e-7)
    static = tl.train_sigmas(1000, 3.0, use_dynamic_shifting=False)
    assert torch.allclose(static[d["idx"]], d["sigmas_static_shift3"].flatten(), rtol=0, atol=1e-7)


def test_kd_oracle_matches_the_reference_loss_loop_verbatim(golden_dir):
    """kd_loop.pt holds the outputs of the reference's LITERAL loss loop (train/train_qwenvl.py:601-620, cut out of train() and
    exec'ed by oracle/make_golden.py::golden_kd_loop).  The oracle's restated loop must reproduce it: value, gradients, and the
    inf/nan guard (layers skipped, names printed)."""
    from oracle import kd_oracle
    d = torch.load(os.path.join(golden_dir, "kd_loop.pt"))
    assert "F.kl_div(F.softmax(normalize(KD_teacher_tensor0[:,i])/temperature0, dim=-1).log()" in d["loop_src"]
    ss = [s.float().clone().requires_grad_(True) for s in d["student"]]
    loss = kd_oracle.kd_loss_stacked(*[t.float() for t in d["teacher"]], *ss)
    grads = torch.autograd.grad(loss, ss)
    assert torch.allclose(loss, d["clean_fp32"]["loss"], rtol=1e-6, atol=1e-7)
    for a, b in zip(grads, d["clean_fp32"]["grads"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)
    # bf16, the reference's own dtype: same ops in the same order -> identical
    lb = kd_oracle.kd_loss_stacked(*d["teacher"], *d["student"])
    assert torch.equal(lb.float(), d["clean_bf16"]["loss"])
    # guard: a NaN / inf in a teacher layer drops exactly that layer's term
    lp = kd_oracle.kd_loss_stacked(*[t.float() for t in d["poisoned_teacher"]], *[s.float() for s in d["student"]])
    assert torch.allclose(lp, d["poisoned_fp32"]["loss"], rtol=1e-6, atol=1e-7)
    assert d["poisoned_fp32"]["printed"] == ["down_feature:2", "down_feature2:5"] and d["clean_fp32"]["printed"] == []
    assert float(d["poisoned_fp32"]["loss"]) < float(d["clean_fp32"]["loss"])


def test_reference_cast_hook_list_registers_on_the_drop_in(golden_dir):
    """The reference's own cast_hook_list (train_qwenvl.py:206-214), run by make_golden on an x2i_b200 transformer through the compat
    shim: one hook per attn module, (img, txt) fan-out for double blocks, tensor for single blocks -- and x2i_b200.kd.cast_hook_list
    produces the same wiring on the same model."""
    from x2i_b200.flux import FluxTransformer2DModel
    from x2i_b200.kd import cast_hook_list
    d = torch.load(os.path.join(golden_dir, "kd_loop.pt"))
    assert d["n_hooks"] == [1] * 7
    assert d["hook_fanout"] == [[0.0, 1.0, 2.0], [100.0, 101.0, 102.0], [200.0, 201.0, 202.0, 203.0]]
    with torch.device("meta"):
        model = FluxTransformer2DModel(num_layers=3, num_single_layers=4, num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32)
    lists = []
    cast_hook_list(model, lists)
    for i, b in enumerate(model.transformer_blocks):
        for h in b.attn._forward_hooks.values():
            h(b.attn, (), (torch.full((1,), float(i)), torch.full((1,), 100.0 + i)))
    for i, b in enumerate(model.single_transformer_blocks):
        for h in b.attn._forward_hooks.values():
            h(b.attn, (), torch.full((1,), 200.0 + i))
    assert [[float(t) for t in lst] for lst in lists] == d["hook_fanout"]


def test_compat_diffusers_shim_surface():
    """x2i_b200.compat.install(): every diffusers name the reference's hot-path files import resolves to the x2i_b200 drop-in."""
    import importlib
    import math
    import sys
    from x2i_b200 import compat, controlnext, flux, pipeline, vae
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] == "diffusers"}
    try:
        compat.install(force=True)
        import diffusers
        from diffusers import AutoencoderKL, FluxPipeline                                            # infer/inference_qwenvl.py:7
        from diffusers.image_processor import VaeImageProcessor                                      # :8
        from diffusers.models.transformers import FluxTransformer2DModel                             # train/train_qwenvl.py:45
        from diffusers.schedulers import FlowMatchEulerDiscreteScheduler                             # :44
        from diffusers.optimization import get_scheduler                                             # :31
        from diffusers.utils.torch_utils import is_compiled_module, randn_tensor                     # :32
        from diffusers.models.attention import FeedForward                                           # lightcontrol_flux.py:24
        from diffusers.models.attention_processor import Attention, FluxAttnProcessor2_0, FusedFluxAttnProcessor2_0  # :25-30
        from diffusers.models.normalization import AdaLayerNormContinuous, AdaLayerNormZero, AdaLayerNormZeroSingle  # :32
        from diffusers.models.embeddings import CombinedTimestepGuidanceTextProjEmbeddings, FluxPosEmbed               # :35
        from diffusers.models.resnet import Downsample2D, ResnetBlock2D                              # :37
        from diffusers.training_utils import compute_density_for_timestep_sampling, compute_loss_weighting_for_sd3  # train_lightcontrol.py:36
        assert FluxTransformer2DModel is flux.FluxTransformer2DModel and FluxPipeline is pipeline.FluxPipeline
        assert AutoencoderKL is vae.AutoencoderKL and VaeImageProcessor is vae.VaeImageProcessor
        assert FlowMatchEulerDiscreteScheduler is pipeline.FlowMatchEulerDiscreteScheduler and Attention is flux.Attention
        assert ResnetBlock2D is controlnext.ResnetBlock2D and Downsample2D is controlnext.Downsample2D
        assert diffusers.utils.check_min_version("0.31.0") is None and not is_compiled_module(torch.nn.Linear(2, 2))
        assert importlib.import_module("diffusers.loaders.lora_pipeline").SD3LoraLoaderMixin is not None
        # the cosine schedule of train_qwenvl.sh (warm-up 100 of 100 000 steps, train_qwenvl.py:476-481)
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
        sch = get_scheduler("cosine", optimizer=opt, num_warmup_steps=100, num_training_steps=100000)
        lrs = []
        for _ in range(151):
            lrs.append(sch.get_last_lr()[0]); opt.step(); sch.step()
        assert lrs[0] == 0.0 and abs(lrs[50] - 0.5e-4) < 1e-12 and abs(lrs[100] - 1e-4) < 1e-12
        assert abs(lrs[150] - 1e-4 * 0.5 * (1 + math.cos(math.pi * 50 / 99900))) < 1e-12
        w = compute_loss_weighting_for_sd3("cosmap", torch.tensor([0.5]))
        assert abs(float(w) - 2 / (math.pi * 0.5)) < 1e-6 and float(compute_loss_weighting_for_sd3("none", torch.tensor([0.3]))) == 1.0
        assert randn_tensor((2, 3), generator=torch.Generator().manual_seed(0)).shape == (2, 3)
        assert compute_density_for_timestep_sampling("logit_normal", 4, 0.0, 1.0).shape == (4,)
    finally:
        compat.uninstall()
        for k in [k for k in sys.modules if k.split(".")[0] == "diffusers"]:
            sys.modules.pop(k)
        sys.modules.update(saved)
