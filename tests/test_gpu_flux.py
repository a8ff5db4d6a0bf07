"""-m gpu: the drop-in FLUX MMDiT / pipeline / projector / KD loss against the fp32 CPU oracle on seeded inputs.

Tolerance: BASELINE.md's 1e-2 relative (Frobenius) for bf16 paths.  The reference's own eager-bf16 path deviates from the
fp32 oracle by a comparable amount (measured below as the yardstick).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import smoke
    return smoke


def _hooks(model):
    from x2i_b200.kd import cast_hook_list
    lists = []
    cast_hook_list(model, lists)
    return lists


@pytest.mark.parametrize("guidance", [False, True])
def test_transformer_step_matches_oracle(env, guidance):
    cfg = env.tiny_config(guidance)
    model, oracle = env.make_pair(cfg, seed=3 + guidance)
    inp = env.make_inputs(cfg, B=2, hl=8, wl=12, S=40, seed=5)
    with torch.no_grad():
        ref = oracle(**env.oracle_inputs(inp), return_dict=False)[0]
        out = model(**env.to_device(inp), return_dict=False)[0]
        # yardstick: the reference's eager bf16 path (oracle modules run in bf16 on the GPU)
        ob = oracle.to("cuda", torch.bfloat16)
        eager = ob(**env.to_device(inp), return_dict=False)[0]
    assert out.shape == ref.shape == (2, 96, 64)
    e_mine, e_eager = env.rel(out, ref), env.rel(eager, ref)
    print(f"rel err vs fp32 oracle: x2i_b200 {e_mine:.4f}, eager-bf16 reference path {e_eager:.4f}")
    assert e_mine < TOL
    assert env.rel(out, eager) < 2 * TOL


def test_forward_hooks_see_reference_tensors(env):
    """B3 contract (train_qwenvl.py:186-214): hooks on blk.attn get (img, txt) / single-attn outputs."""
    cfg = env.tiny_config(True)
    model, oracle = env.make_pair(cfg, seed=7)
    inp = env.make_inputs(cfg, B=2, hl=8, wl=8, S=24, seed=8)
    hm, ho = _hooks(model), _hooks(oracle)
    with torch.no_grad():
        ref = oracle(**env.oracle_inputs(inp), return_dict=False)[0]
        out = model(**env.to_device(inp), return_dict=False)[0]
    assert env.rel(out, ref) < TOL
    assert [len(x) for x in hm] == [2, 2, 3]
    for gm, go in zip(hm, ho):
        for a, b in zip(gm, go):
            assert a.shape == b.shape
            assert env.rel(a, b) < TOL


def test_plugin_processor_path_matches_fused_path(env):
    """B4 contract: a user processor invoked with the diffusers protocol; result must equal the fused default."""
    from x2i_b200.flux import FluxAttnProcessor2_0

    class Wrapped:  # a plug-in that simply defers to the stock maths through the public protocol
        calls = 0

        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, image_rotary_emb=None):
            Wrapped.calls += 1
            return FluxAttnProcessor2_0()(attn, hidden_states, encoder_hidden_states, attention_mask, image_rotary_emb)

    cfg = env.tiny_config(False)
    model, oracle = env.make_pair(cfg, seed=9)
    inp = env.to_device(env.make_inputs(cfg, B=1, hl=8, wl=8, S=16, seed=10))
    with torch.no_grad():
        a = model(**inp, return_dict=False)[0].clone()
        model.set_attn_processor(Wrapped())
        b = model(**inp, return_dict=False)[0]
    assert Wrapped.calls == 5
    assert env.rel(b, a) < 5e-3


def test_pipeline_four_step_sampling_matches_oracle(env):
    from oracle import flux_oracle as fo
    from x2i_b200.pipeline import FlowMatchEulerDiscreteScheduler, FluxPipeline
    cfg = env.tiny_config(False)
    model, oracle = env.make_pair(cfg, seed=11)
    pipe = FluxPipeline(scheduler=FlowMatchEulerDiscreteScheduler(shift=1.0), transformer=model)
    g = torch.Generator().manual_seed(12)
    bf = lambda t: t.to(torch.bfloat16)  # noqa: E731
    prompt, pooled = bf(torch.randn(2, 24, 64, generator=g)), bf(torch.randn(2, 32, generator=g))
    lat = bf(torch.randn(2, 16, 16, 16, generator=g))  # [B,16,h,w] -> packed [B,64,64]
    packed = fo.pack_latents(lat)
    out = pipe(prompt_embeds=prompt.cuda(), pooled_prompt_embeds=pooled.cuda(), num_inference_steps=4, guidance_scale=3.5,
               height=128, width=128, output_type="latent", latents=packed.cuda()).images
    ref = fo.denoise(oracle, packed.float(), prompt.float(), pooled.float(), 16, 16, 4, emulate_bf16_time=True)
    assert out.shape == (2, 64, 64)
    assert env.rel(out, ref) < TOL
    # generator path: shapes and determinism
    a = pipe(prompt_embeds=prompt.cuda(), pooled_prompt_embeds=pooled.cuda(), num_inference_steps=2, height=128, width=128,
             output_type="latent", generator=torch.Generator("cpu").manual_seed(3)).images
    b = pipe(prompt_embeds=prompt.cuda(), pooled_prompt_embeds=pooled.cuda(), num_inference_steps=2, height=128, width=128,
             output_type="latent", generator=torch.Generator("cpu").manual_seed(3)).images
    assert torch.equal(a, b)
    assert FluxPipeline._unpack_latents(a, 128, 128, 16).shape == (2, 16, 16, 16)
    # caller-supplied latents already in bf16 on the device are NOT overwritten by the in-place Euler updates
    mine = packed.cuda()
    keep = mine.clone()
    o1 = pipe(prompt_embeds=prompt.cuda(), pooled_prompt_embeds=pooled.cuda(), num_inference_steps=4, height=128, width=128,
              output_type="latent", latents=mine).images
    assert torch.equal(mine, keep) and torch.equal(o1, out)
    # every step's AdaLN modulation from ONE pass over the modulation weights (default) == the per-step GEMV, bit for bit; also on the
    # eager path and with the guidance embedding (dev configuration)
    for cfg2, graph in ((cfg, True), (cfg, False), (env.tiny_config(True), True)):
        m2 = model if cfg2 is cfg else env.make_pair(cfg2, seed=13)[0]
        p2 = FluxPipeline(scheduler=FlowMatchEulerDiscreteScheduler(shift=1.0), transformer=m2)
        m2.use_cuda_graph = graph
        kw = dict(prompt_embeds=prompt.cuda(), pooled_prompt_embeds=pooled.cuda(), num_inference_steps=3, height=128, width=128,
                  output_type="latent", latents=packed.cuda(), guidance_scale=2.5)
        try:
            hoisted = p2(**kw).images
            p2.hoist_modulation = False
            plain = p2(**kw).images
        finally:
            m2.use_cuda_graph = True
        assert torch.equal(hoisted, plain)
    mods = model.precompute_modulation(torch.tensor([1.0, 0.5], device="cuda"), pooled.cuda())
    assert mods.shape[:2] == (2, 2) and mods.dtype == torch.bfloat16


@pytest.mark.parametrize("S,hl,wl,outliers", [(203, 16, 24, False), (512, 64, 64, False), (512, 64, 64, True)])
def test_full_width_blocks_match_oracle(env, S, hl, wl, outliers):
    """One double + one single block at the real width (D=3072, 24 heads): a ragged sequence (S_txt=203, MiniCPM's unpadded
    prompts) and BASELINE's FULL size (512 text + 4096 latent tokens, 1024 px), the latter also with outlier channels
    (4 channels of both streams x40, SURVEY.md 8(d): real FLUX activations carry such channels).  fp32 oracle on the GPU."""
    from oracle import flux_oracle as fo
    from x2i_b200 import flux as xf
    torch.manual_seed(0)
    D, H, B = 3072, 24, 1
    L_img = hl * wl
    od = fo.FluxTransformerBlock(D, H, 128).eval()
    os_ = fo.FluxSingleTransformerBlock(D, H, 128).eval()
    for m in (od, os_):
        fo.init_synthetic_(m, seed=21, std=0.02)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(p.bfloat16().float())
    md = xf.FluxTransformerBlock(D, H, 128); md.load_state_dict(od.state_dict()); md = md.to("cuda", torch.bfloat16)
    ms = xf.FluxSingleTransformerBlock(D, H, 128); ms.load_state_dict(os_.state_dict()); ms = ms.to("cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(22)
    bf = lambda t: t.bfloat16().float()  # noqa: E731
    x, c, temb = torch.randn(B, L_img, D, generator=g), torch.randn(B, S, D, generator=g), bf(torch.randn(B, D, generator=g))
    if outliers:
        ch = torch.randperm(D, generator=g)[:4]
        x[..., ch] *= 40.0
        c[..., ch] *= 40.0
    x, c = bf(x), bf(c)
    ids = torch.cat([torch.zeros(S, 3), fo.prepare_latent_image_ids(2 * hl, 2 * wl)])
    rope = fo.rope_table(ids)
    dev = lambda t: t.to("cuda", torch.bfloat16)  # noqa: E731
    rope_dev = (rope[0].cuda(), rope[1].cuda())
    with torch.no_grad():
        od, os_ = od.cuda(), os_.cuda()  # fp32 oracle on the GPU (size)
        rc, rx = od(x.cuda(), c.cuda(), temb.cuda(), rope_dev)
        rh = os_(torch.cat([rc, rx], 1), temb.cuda(), rope_dev)
        c2, x2 = md(dev(x), dev(c), dev(temb), image_rotary_emb=rope_dev)
        h2 = ms(torch.cat([c2, x2], 1).contiguous(), dev(temb), image_rotary_emb=rope_dev)
    assert env.rel(x2, rx) < TOL and env.rel(c2, rc) < TOL
    assert env.rel(h2, rh) < TOL


def test_projector_matches_oracle_and_reference_golden(env, golden_dir):
    from oracle import proj_oracle
    from oracle.make_golden import synth_state
    from x2i_b200 import proj as xproj
    # small, all three branches, weights and inputs of the fixture minted from the reference's utils/proj.py
    gsmall = torch.load(os.path.join(golden_dir, "proj_small.pt"), weights_only=False)
    # (H=64 fixture is below the kernels' GEMM granularity for N=24; use config-1 instead for the reference pin)
    g1 = torch.load(os.path.join(golden_dir, "proj_c1.pt"), weights_only=False)
    p = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True)
    sd = synth_state(p, 21, std=0.02)
    p.load_state_dict(sd)
    p = p.to("cuda", torch.bfloat16)
    x = torch.randn(1, 37, 77, 2048, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        pooled, seq = p(x.cuda())
    assert pooled.shape == (1, 768) and seq.shape == (1, 77, 4096)
    assert env.rel(pooled, g1["pooled"]) < 2e-2           # reference fp32 weights vs bf16 weights + bf16 activations
    assert env.rel(seq[:, ::7, ::13], g1["seq_sub"]) < 2e-2
    # same bf16-rounded weights in the fp32 oracle -> tight comparison, all branches, 7B shape with ragged S
    for kind, C, kw in (("qwen7b", 29, dict(use_scale=False, use_cnn=True)), ("internvl1b", 25, dict(use_scale=True, use_cnn=True)),
                        ("qwen3b", 37, dict(use_scale=False, use_cnn=False))):
        o = proj_oracle.create_proj(kind, C, **kw)
        sd = {k: v.bfloat16().float() for k, v in synth_state(o, 5, std=0.03).items()}
        o.load_state_dict(sd)
        m = getattr(xproj, {"qwen7b": "create_proj3_qwen7b", "internvl1b": "create_proj_internvl1b", "qwen3b": "create_proj3_qwen3b"}[kind])(
            C, use_t5=False, **kw)
        m.load_state_dict(sd)
        m = m.to("cuda", torch.bfloat16)
        H = o.mlp.layernorm.normalized_shape[0]
        x = torch.randn(2, C, 45, H, generator=torch.Generator().manual_seed(6)).bfloat16()
        with torch.no_grad():
            rp, rs = o(x.float())
            mp_, ms_ = m(x.cuda())
        assert env.rel(ms_, rs) < TOL and env.rel(mp_, rp) < TOL, kind


def test_kd_loss_module_matches_oracle_stacked(env):
    from oracle import kd_oracle
    from x2i_b200.kd import attention_distillation_loss
    g = torch.Generator().manual_seed(31)
    B, D = 2, 3072
    shapes = [(B, 3, 64, D), (B, 3, 24, D), (B, 4, 88, D)]
    T = [torch.randn(s, generator=g).bfloat16() for s in shapes]
    S = [(t.float() + 0.5 * torch.randn(s, generator=g)).bfloat16() for t, s in zip(T, shapes)]
    S_ref = [s.float().requires_grad_(True) for s in S]
    ref = kd_oracle.kd_loss_stacked(*[t.float() for t in T], *S_ref)
    ref.backward()
    S_dev = [s.cuda().requires_grad_(True) for s in S]
    loss = attention_distillation_loss([t.cuda() for t in T], S_dev, temperature=3.0, verbose=False)
    loss.backward()
    assert abs(float(loss) - float(ref)) / float(ref) < 2e-3
    for a, b in zip(S_dev, S_ref):
        assert a.grad.shape == b.grad.shape
        assert env.rel(a.grad, b.grad) < TOL


def test_resampler_matches_reference_golden_and_oracle(env, golden_dir):
    from oracle import resampler_oracle as ro
    from oracle.make_golden import synth_state
    from x2i_b200.resampler import Resampler
    # (1) the reference's own output (fixture minted from minicpm/resampler.py), ragged tgt_sizes incl. a 5x1 grid
    g = torch.load(os.path.join(golden_dir, "resampler_small.pt"), weights_only=False)
    m = Resampler(num_queries=8, embed_dim=256, num_heads=2, kv_dim=48, adaptive=True, max_size=(6, 7))
    m.load_state_dict(g["state"])
    m = m.to("cuda", torch.bfloat16)
    y = m(g["x"].cuda(), g["tgt_sizes"].cuda())
    assert y.shape == g["out"].shape
    assert env.rel(y, g["out"]) < 2e-2   # reference fp32 weights/activations vs bf16 everywhere
    # (2) MiniCPM-o-2.6 dimensions (64 queries, 3584 = 28 x 128, kv 1152), same bf16-rounded weights in the fp32 oracle
    o = ro.Resampler(num_queries=64, embed_dim=3584, num_heads=28, kv_dim=1152, adaptive=True).eval()
    sd = {k: v.bfloat16().float() for k, v in synth_state(o, 61, std=0.02).items()}
    sd["query"] = (torch.randn(64, 3584, generator=torch.Generator().manual_seed(62)) * 0.5).bfloat16().float()
    o.load_state_dict(sd)
    m = Resampler(num_queries=64, embed_dim=3584, num_heads=28, kv_dim=1152, adaptive=True)
    m.load_state_dict(sd)
    m = m.to("cuda", torch.bfloat16)
    tgt = torch.tensor([[20, 30], [24, 24], [7, 72]])  # 72 > default 70-wide table: exercises _adjust_pos_cache
    x = torch.randn(3, 600, 1152, generator=torch.Generator().manual_seed(63)).bfloat16()
    with torch.no_grad():
        ref = o.cuda()(x.float().cuda(), tgt.cuda())
    y = m(x.cuda(), tgt.cuda())
    assert y.shape == (3, 64, 3584)
    assert env.rel(y, ref) < TOL


def test_mismatched_position_ids_fail_loudly(env):
    """ids shorter than the token sequence would index the RoPE table out of bounds; the reference fails in apply_rotary_emb."""
    from x2i_b200._lib import X2IError
    cfg = env.tiny_config(False)
    model, _ = env.make_pair(cfg, seed=3)
    inp = env.to_device(env.make_inputs(cfg, B=1, hl=8, wl=8, S=16, seed=10))
    inp["img_ids"] = inp["img_ids"][: inp["img_ids"].shape[0] // 2].contiguous()
    with pytest.raises(X2IError):
        with torch.no_grad():
            model(**inp, return_dict=False)


def test_kd_loss_matches_the_reference_loss_loop_fixture(env, golden_dir):
    """x2i_b200.kd.attention_distillation_loss against tests/golden/kd_loop.pt: the outputs of the reference's LITERAL loss loop
    (train/train_qwenvl.py:601-620, exec'ed unchanged by oracle/make_golden.py::golden_kd_loop) in fp32; yardstick = the same lines
    in the reference's own dtype (bf16).  Also the inf/nan guard: the poisoned layers are skipped and named like the reference."""
    import contextlib
    import io
    from parity import check
    from x2i_b200 import kd
    d = torch.load(os.path.join(golden_dir, "kd_loop.pt"))
    t = [x.cuda() for x in d["teacher"]]
    s = [x.cuda().clone().requires_grad_(True) for x in d["student"]]
    loss = kd.attention_distillation_loss(t, s, temperature=3.0, verbose=False)
    grads = torch.autograd.grad(loss, s)
    ref, eag = d["clean_fp32"], d["clean_bf16"]
    check("KD loss vs the reference's literal loop (kd_loop.pt): loss", abs(float(loss) - float(ref["loss"])) / float(ref["loss"]),
          abs(float(eag["loss"]) - float(ref["loss"])) / float(ref["loss"]))
    for i, (a, b, e) in enumerate(zip(grads, ref["grads"], eag["grads"])):
        check(f"KD loss vs the reference's literal loop (kd_loop.pt): d student group {i}", env.rel(a, b), env.rel(e, b))
    # hook LISTS instead of stacked tensors: same value
    l2 = kd.attention_distillation_loss([list(x.unbind(1)) for x in t], [list(x.detach().unbind(1)) for x in s], verbose=False)
    assert abs(float(l2) - float(loss)) <= 1e-5 * abs(float(loss))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        lp = kd.attention_distillation_loss([x.cuda() for x in d["poisoned_teacher"]], [x.cuda() for x in d["student"]], verbose=True)
    assert buf.getvalue().split() == d["poisoned_fp32"]["printed"] == ["down_feature:2", "down_feature2:5"]
    assert abs(float(lp) - float(d["poisoned_fp32"]["loss"])) / float(d["poisoned_fp32"]["loss"]) < 1e-2
