"""-m gpu: the differentiable student pass (forward in saving mode + hand-written backward through every block) against
fp32 autograd of the oracle on the same seeded inputs.  The yardstick printed next to each error is the reference's own
eager-bf16 path (oracle modules in bf16 on the GPU, PyTorch autograd) measured against the same fp32 ground truth."""
import pytest
import torch

from parity import check

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


def _probe_loss(out, hooks, seed, device):
    """A generic scalar of the model output and every hooked tensor (fixed random probes) so that all gradient paths are hit."""
    g = torch.Generator().manual_seed(seed)
    loss = (out.float() * torch.randn(out.shape, generator=g).to(device)).sum()
    for lst in hooks:
        for t in lst:
            loss = loss + (t.float() * torch.randn(t.shape, generator=g).to(device)).sum()
    return loss


def _run(model, inp, device, seed, keys=("encoder_hidden_states", "pooled_projections")):
    inp = dict(inp)
    for k in keys:
        inp[k] = inp[k].detach().clone().requires_grad_(True)
    hooks = _hooks(model)
    out = model(**inp, return_dict=False)[0]
    loss = _probe_loss(out, hooks, seed, device)
    grads = torch.autograd.grad(loss, [inp[k] for k in keys])
    for m in model.modules():
        m._forward_hooks.clear()
    return out.detach(), [h for lst in hooks for h in lst], grads


@pytest.mark.parametrize("guidance,B,hl,wl,S", [(True, 2, 8, 8, 24), (False, 1, 8, 12, 40)])
def test_student_backward_matches_oracle_autograd(env, guidance, B, hl, wl, S):
    cfg = env.tiny_config(guidance)
    model, oracle = env.make_pair(cfg, seed=21)
    inp = env.make_inputs(cfg, B=B, hl=hl, wl=wl, S=S, seed=22)
    o_ref, h_ref, g_ref = _run(oracle, env.oracle_inputs(inp), "cpu", 5)
    o_mine, h_mine, g_mine = _run(model, env.to_device(inp), "cuda", 5)
    oracle_bf = oracle.to("cuda", torch.bfloat16)
    _, _, g_eager = _run(oracle_bf, env.to_device(inp), "cuda", 5)
    assert env.rel(o_mine, o_ref) < TOL
    for a, b in zip(h_mine, h_ref):
        assert env.rel(a, b) < TOL
    for name, a, b, e in zip(("d encoder_hidden_states", "d pooled_projections"), g_mine, g_ref, g_eager):
        assert a.shape == b.shape and a.dtype == torch.bfloat16
        check(f"tiny student backward (guidance={guidance}, B={B}): {name}", env.rel(a, b), env.rel(e, b))


def test_gradient_checkpointing_recomputes_to_identical_gradients(env):
    """enable_gradient_checkpointing() (train_lightcontrol.py:666; lightcontrol_flux.py:475-494,:513-531): the differentiable forward
    keeps only block inputs and the backward re-runs each block's forward kernels -- same kernels on the same data, so outputs,
    hooks and gradients are bit-identical to the saving mode, with less memory held between forward and backward."""
    cfg = env.tiny_config(True)
    model, _ = env.make_pair(cfg, seed=61)
    inp = env.to_device(env.make_inputs(cfg, B=2, hl=8, wl=12, S=24, seed=62))
    o0, h0, g0 = _run(model, inp, "cuda", 9)
    model.enable_gradient_checkpointing()
    assert model.gradient_checkpointing
    torch.cuda.reset_peak_memory_stats()
    o1, h1, g1 = _run(model, inp, "cuda", 9)
    model.disable_gradient_checkpointing()
    assert torch.equal(o0, o1)
    assert all(torch.equal(a, b) for a, b in zip(h0, h1))
    assert all(torch.equal(a, b) for a, b in zip(g0, g1))


def test_hidden_states_gradient_and_no_hooks(env):
    cfg = env.tiny_config(False)
    model, oracle = env.make_pair(cfg, seed=23)
    inp = env.make_inputs(cfg, B=1, hl=8, wl=8, S=16, seed=24)
    keys = ("hidden_states", "encoder_hidden_states", "pooled_projections")

    def run(m, i, dev):
        i = dict(i)
        for k in keys:
            i[k] = i[k].detach().clone().requires_grad_(True)
        out = m(**i, return_dict=False)[0]
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).to(dev)
        return torch.autograd.grad((out.float() * w).sum(), [i[k] for k in keys])

    g_ref = run(oracle, env.oracle_inputs(inp), "cpu")
    g_mine = run(model, env.to_device(inp), "cuda")
    g_eag = run(oracle.to("cuda", torch.bfloat16), env.to_device(inp), "cuda")
    for k, a, b, e in zip(keys, g_mine, g_ref, g_eag):
        check(f"tiny student backward without hooks: d {k}", env.rel(a, b), env.rel(e, b))


def test_kd_training_step_matches_oracle(env):
    """Teacher pass (no grad) + student pass + the reference's KD loss + backward: loss and gradients vs the fp32 oracle."""
    from oracle import kd_oracle
    from x2i_b200 import kd
    cfg = env.tiny_config(True)
    model, oracle = env.make_pair(cfg, seed=25)
    t_inp = env.make_inputs(cfg, B=2, hl=8, wl=8, S=24, seed=26)       # teacher conditioning (T5/CLIP stand-in)
    s_inp = dict(t_inp)
    g = torch.Generator().manual_seed(27)
    s_inp["encoder_hidden_states"] = (t_inp["encoder_hidden_states"] + 0.3 * torch.randn(t_inp["encoder_hidden_states"].shape, generator=g)).to(torch.bfloat16).float()
    s_inp["pooled_projections"] = (t_inp["pooled_projections"] + 0.3 * torch.randn(t_inp["pooled_projections"].shape, generator=g)).to(torch.bfloat16).float()

    def step(m, ti, si, loss_fn):
        th = _hooks(m)
        with torch.no_grad():
            m(**ti, return_dict=False)
        for mod in m.modules():
            mod._forward_hooks.clear()
        si = dict(si)
        for k in ("encoder_hidden_states", "pooled_projections"):
            si[k] = si[k].detach().clone().requires_grad_(True)
        sh = _hooks(m)
        m(**si, return_dict=False)
        for mod in m.modules():
            mod._forward_hooks.clear()
        loss = loss_fn([torch.stack(x, 1) for x in th], [torch.stack(x, 1) for x in sh])
        grads = torch.autograd.grad(loss, [si["encoder_hidden_states"], si["pooled_projections"]])
        return float(loss.detach()), grads

    l_ref, g_ref = step(oracle, env.oracle_inputs(t_inp), env.oracle_inputs(s_inp),
                        lambda t, s: kd_oracle.kd_loss_stacked(*t, *s))
    l_mine, g_mine = step(model, env.to_device(t_inp), env.to_device(s_inp),
                          lambda t, s: kd.attention_distillation_loss(t, s, verbose=False))
    # yardstick: the reference's own path -- the oracle modules in bf16 with torch eager ops and the reference's literal loss
    l_eag, g_eag = step(oracle.to("cuda", torch.bfloat16), env.to_device(t_inp), env.to_device(s_inp),
                        lambda t, s: kd_oracle.kd_loss_stacked(*t, *s))
    print(f"KD loss: x2i_b200 {l_mine:.6f}  fp32 oracle {l_ref:.6f}  eager bf16 {l_eag:.6f}")
    # the loss is a small difference of near-equal distributions: bf16 paths deviate by more than 1e-2 from fp32
    check("tiny KD train step: loss", abs(l_mine - l_ref) / abs(l_ref), abs(l_eag - l_ref) / abs(l_ref))
    for name, a, b, e in zip(("d encoder_hidden_states", "d pooled_projections"), g_mine, g_ref, g_eag):
        check(f"tiny KD train step: {name}", env.rel(a, b), env.rel(e, b))


@pytest.mark.parametrize("kind,use_scale,use_cnn,C,S,H", [("qwen3b", False, True, 5, 24, 256), ("qwen7b", True, False, 4, 20, 256),
                                                           ("internvl4b", False, False, 3, 16, 128), ("qwen3b", False, True, 37, 77, 2048)])
def test_projector_backward_matches_reference_autograd(env, kind, use_scale, use_cnn, C, S, H):
    """Every projector parameter gradient vs fp32 autograd of the projector oracle (pinned to utils/proj.py)."""
    from oracle import proj_oracle
    from x2i_b200 import proj as xproj
    torch.manual_seed(31)
    B = 2 if H < 2048 else 1
    o = proj_oracle.Proj7Exp(in_channels=C, input_dim=H, use_scale=use_scale, use_cnn=use_cnn)
    with torch.no_grad():
        for p_ in o.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    m = xproj.Proj7Exp(in_channels=C, input_dim=H, use_t5=False, use_scale=use_scale, use_cnn=use_cnn)
    m.load_state_dict(o.state_dict())
    m = m.to("cuda", torch.bfloat16)
    x = torch.randn(B, C, S, H).to(torch.bfloat16)
    w1, w2 = torch.randn(B, 768), torch.randn(B, S, 4096)
    p1, p2 = o(x.float())
    ((p1 * w1).sum() + (p2 * w2).sum()).backward()
    q1, q2 = m(x.cuda())
    assert env.rel(q1, p1) < TOL and env.rel(q2, p2) < TOL
    ((q1.float() * w1.cuda()).sum() + (q2.float() * w2.cuda()).sum()).backward()
    ref = dict(o.named_parameters())
    # yardstick: the reference projector itself in bf16 under torch autograd (what train_qwenvl.py:399,:576,:625 runs)
    import copy
    ob = copy.deepcopy(o).to("cuda", torch.bfloat16)
    ob.zero_grad()
    e1, e2 = ob(x.cuda())
    ((e1.float() * w1.cuda()).sum() + (e2.float() * w2.cuda()).sum()).backward()
    eag = dict(ob.named_parameters())
    for name, p_ in m.named_parameters():
        assert p_.grad is not None, name
        if name == "conv.bias":
            # a constant added in front of a LayerNorm has an exactly-zero gradient; both sides hold rounding noise only
            assert float(p_.grad.float().abs().max()) < 1e-2 * float(ref["conv.weight"].grad.norm())
            continue
        check(f"projector backward {kind} C={C} S={S} H={H}: {name}", env.rel(p_.grad, ref[name].grad),
              env.rel(eag[name].grad, ref[name].grad))


def test_distill_step_end_to_end(env):
    """distill_step: teacher + student + KD + backward + clip + AdamW on a tiny FLUX; list-based and stacked losses agree and
    the projector gradients match the oracle pipeline (projector oracle -> FLUX oracle -> KD oracle, fp32 autograd)."""
    from oracle import kd_oracle, proj_oracle
    from x2i_b200 import proj as xproj, train
    cfg = dict(env.tiny_config(True), joint_attention_dim=4096, pooled_projection_dim=768)
    model, oracle = env.make_pair(cfg, seed=41)
    torch.manual_seed(42)
    C, S, H, B, hl, wl = 4, 24, 128, 2, 8, 8
    po = proj_oracle.Proj7Exp(in_channels=C, input_dim=H, use_scale=False, use_cnn=True)
    with torch.no_grad():
        for p_ in po.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    pm = xproj.Proj7Exp(in_channels=C, input_dim=H, use_t5=False, use_scale=False, use_cnn=True)
    pm.load_state_dict(po.state_dict())
    pm = pm.to("cuda", torch.bfloat16)
    bf = lambda t: t.to(torch.bfloat16)  # noqa: E731
    batch = dict(latents=bf(torch.randn(B, hl * wl, 64)), timestep=bf(torch.full((B,), 1000.0)),
                 text_embeddings=bf(torch.randn(B, C, S, H)), prompt_embeds_t5=bf(torch.randn(B, S, 4096) * 0.2),
                 pooled_clip=bf(torch.randn(B, 768) * 0.2))
    cb = {k: v.cuda() for k, v in batch.items()}
    loss_a = train.distill_step(pm, model, cb, optimizer=None, height=2 * hl, width=2 * wl)
    grads_a = {n: p_.grad.clone() for n, p_ in pm.named_parameters()}
    pm.zero_grad()
    loss_b = train.distill_step(pm, model, cb, optimizer=None, height=2 * hl, width=2 * wl, stacked=True)
    assert abs(float(loss_a) - float(loss_b)) / float(loss_b) < 1e-3
    for n, p_ in pm.named_parameters():
        if n != "conv.bias":
            assert env.rel(p_.grad, grads_a[n]) < 2e-3, n
    # oracle pipeline
    from x2i_b200.kd import cast_hook_list
    import oracle.flux_oracle as fo
    common = dict(hidden_states=batch["latents"].float(), timestep=torch.full((B,), 1.0), txt_ids=torch.zeros(S, 3),
                  img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), guidance=(bf(torch.full((B,), 3.5)) * 1000).float() / 1000)
    th = []
    cast_hook_list(oracle, th)
    with torch.no_grad():
        oracle(encoder_hidden_states=batch["prompt_embeds_t5"].float(), pooled_projections=batch["pooled_clip"].float(), return_dict=False, **common)
    for mod in oracle.modules():
        mod._forward_hooks.clear()
    sh = []
    cast_hook_list(oracle, sh)
    a, e = po(batch["text_embeddings"].float())
    oracle(encoder_hidden_states=e, pooled_projections=a, return_dict=False, **common)
    for mod in oracle.modules():
        mod._forward_hooks.clear()
    loss_o = kd_oracle.kd_loss_stacked(*[torch.stack(x, 1) for x in th], *[torch.stack(x, 1) for x in sh])
    loss_o.backward()
    # yardstick: the same pipeline on the reference's own path (projector + FLUX oracle in bf16, torch eager autograd)
    import copy
    pb, ob = copy.deepcopy(po).to("cuda", torch.bfloat16), oracle.to("cuda", torch.bfloat16)
    pb.zero_grad()
    cbf = {k: (v.cuda() if k in ("txt_ids", "img_ids") else v.to("cuda", torch.bfloat16)) for k, v in common.items()}
    cbf["timestep"], cbf["guidance"] = torch.full((B,), 1.0, device="cuda"), torch.full((B,), 3.5, device="cuda")
    th2 = []
    cast_hook_list(ob, th2)
    with torch.no_grad():
        ob(encoder_hidden_states=cb["prompt_embeds_t5"], pooled_projections=cb["pooled_clip"], return_dict=False, **cbf)
    for mod in ob.modules():
        mod._forward_hooks.clear()
    sh2 = []
    cast_hook_list(ob, sh2)
    a2, e2 = pb(cb["text_embeddings"])
    ob(encoder_hidden_states=e2, pooled_projections=a2, return_dict=False, **cbf)
    for mod in ob.modules():
        mod._forward_hooks.clear()
    loss_e = kd_oracle.kd_loss_stacked(*[torch.stack(x, 1) for x in th2], *[torch.stack(x, 1) for x in sh2])
    loss_e.backward()
    eag = dict(pb.named_parameters())
    print(f"distill loss: x2i_b200 {float(loss_a):.6f}  oracle {float(loss_o):.6f}  eager bf16 {float(loss_e):.6f}")
    check("tiny distill step: loss", abs(float(loss_a) - float(loss_o)) / float(loss_o), abs(float(loss_e) - float(loss_o)) / float(loss_o))
    for n, p_ in po.named_parameters():
        if n == "conv.bias":
            continue  # exactly zero in exact arithmetic (constant in front of a LayerNorm)
        check(f"tiny distill step: grad {n}", env.rel(grads_a[n], p_.grad), env.rel(eag[n].grad, p_.grad))
    # optimizer path: parameters move, loss stays finite
    opt = torch.optim.AdamW(pm.parameters(), lr=1e-4, fused=True)
    before = pm.mlp.projector[0].weight.detach().clone()
    l2 = train.distill_step(pm, model, cb, optimizer=opt, height=2 * hl, width=2 * wl)
    assert torch.isfinite(l2) and not torch.equal(before, pm.mlp.projector[0].weight)
