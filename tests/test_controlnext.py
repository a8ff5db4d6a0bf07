"""ControlNeXt (LightControl editing branch, lightcontrol_flux.py:504-507, :575-749).
CPU: the oracle restatement against the fixture minted from the reference's own ControlNeXtModel / FluxTransformer2DModel
classes (oracle/make_golden.py controlnext).  -m gpu: the sm_100a kernels and the drop-in module against the oracle."""
import os

import pytest
import torch

from parity import check, record

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "controlnext.pt")
TOL = 1e-2


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _oracle_net(seed, std):
    from oracle import controlnext_oracle as co
    from oracle.make_golden import synth_state
    net = co.ControlNeXtModel().eval()
    net.load_state_dict(synth_state(net, seed, std=std))
    return net


def test_oracle_controlnext_matches_reference_class():
    d = torch.load(GOLD)
    net = _oracle_net(d["seed"], d["std"])
    assert sorted(net.state_dict().keys()) == d["keys"] and sum(p.numel() for p in net.parameters()) == d["n_params"]
    with torch.no_grad():
        o = net(d["hint"], d["timestep"])
    assert o["scale"] == d["scale"] == 1.0
    assert o["out"].shape == d["out"].shape == (2, 3072, 4, 6)
    assert torch.allclose(o["out"], d["out"], atol=1e-5, rtol=1e-5)


def test_oracle_injection_matches_reference_transformer():
    """The reference FluxTransformer2DModel with control_nets (2 double blocks, 1 net -> `index_block < len(control_nets)`)."""
    from oracle import flux_oracle as fo
    from oracle.make_golden import synth_state
    d = torch.load(GOLD)
    t = d["transformer"]
    model = fo.FluxTransformer2DModel(**t["cfg"]).eval()
    model.load_state_dict(synth_state(model, t["seeds"]["transformer"], std=t["seeds"]["std_t"]))
    nets = [_oracle_net(s, t["seeds"]["std_n"]) for s in t["seeds"]["nets"]]
    with torch.no_grad():
        y = model(**t["inputs"], guided_hint=d["hint"], control_nets=nets, return_dict=False)[0]
        y0 = model(**t["inputs"], return_dict=False)[0]
    assert torch.allclose(y, t["output"], atol=2e-4, rtol=2e-4)
    assert rel(y0, t["output"]) > 1e-2  # the injection is not a no-op on this fixture


# ------------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import ops
    return ops


def rn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


@gpu
@pytest.mark.parametrize("N,H,W,Cin,Cout,k,stride,pad", [(2, 24, 40, 64, 64, 3, 1, 1), (1, 32, 32, 64, 128, 3, 1, 1), (2, 20, 36, 128, 128, 3, 2, 1),
                                                         (1, 16, 48, 128, 256, 1, 1, 0), (2, 12, 20, 256, 256, 3, 1, 1),
                                                         (1, 8, 12, 256, 3072, 2, 2, 0), (1, 64, 64, 256, 256, 3, 2, 1)])
def test_conv2d_implicit_gemm(ops, N, H, W, Cin, Cout, k, stride, pad):
    import torch.nn.functional as F
    x = rn(N, H, W, Cin, seed=1)
    w = rn(Cout, Cin, k, k, seed=2, scale=0.05)
    b = rn(Cout, seed=3, scale=0.2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    out = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), b, k, k, stride=stride, pad=pad)
    assert out.shape == ref.shape
    assert rel(out, ref) < 4e-3
    # fused epilogue: + per-image channel vector, ReLU, + residual
    rv, res = rn(N, Cout, seed=4), rn(*ref.shape, seed=5)
    out2 = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), b, k, k, stride=stride, pad=pad, rowvec=rv, residual=res, relu=True)
    ref2 = torch.relu(ref + rv.float()[:, None, None, :]) + res.float()
    assert rel(out2, ref2) < 4e-3
    # in place on the residual (the injection form)
    acc = res.clone()
    ops.conv2d_nhwc(x, ops.pack_conv_weight(w), b, k, k, stride=stride, pad=pad, residual=acc, out=acc)
    assert rel(acc, ref + res.float()) < 4e-3


@gpu
def test_conv_first_and_groupnorm(ops):
    import torch.nn.functional as F
    x = rn(2, 3, 40, 56, seed=6)
    w = torch.randn(64, 3, 3, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) * 0.2
    b = torch.randn(64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8)) * 0.1
    out = ops.conv_first(x, w, b)
    ref = F.conv2d(x.float(), w, b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape == (2, 20, 28, 64) and rel(out, ref) < 4e-3
    for C, G, act in ((64, 2, 1), (128, 4, 2), (256, 8, 0), (128, 2, 1)):
        y = rn(2, 36, 52, C, seed=9) * 2 + 0.5
        ga, be, res = (1 + 0.2 * rn(C, seed=10).float()).to(torch.bfloat16), rn(C, seed=11, scale=0.3), rn(2, 36, 52, C, seed=12)
        r = F.group_norm(y.float().permute(0, 3, 1, 2), G, ga.float(), be.float(), eps=1e-5)
        r = (torch.relu(r) if act == 1 else (F.silu(r) if act == 2 else r)).permute(0, 2, 3, 1)
        o = ops.groupnorm_nhwc(y, ga, be, G, 1e-5, act=act)
        assert rel(o, r) < 4e-3
        o2 = ops.groupnorm_nhwc(y, ga, be, G, 1e-5, act=act, residual=res)
        assert rel(o2, r + res.float()) < 4e-3


@gpu
def test_controlnext_model_matches_oracle_and_reference_fixture(ops):
    from x2i_b200.controlnext import ControlNeXtModel
    d = torch.load(GOLD)
    net_o = _oracle_net(d["seed"], d["std"])
    with torch.no_grad():
        for p_ in net_o.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    net = ControlNeXtModel().eval()
    net.load_state_dict(net_o.state_dict())
    net = net.to("cuda", torch.bfloat16)
    hint = d["hint"].to(torch.bfloat16)
    t = (d["timestep"].to(torch.bfloat16))
    with torch.no_grad():
        ref = net_o(hint.float(), t.float())["out"]
        out = net(hint.cuda(), t.cuda())
        eager = net_o.to("cuda", torch.bfloat16)(hint.cuda(), t.cuda())["out"]
    assert out["scale"] == 1.0 and out["out"].shape == ref.shape == (2, 3072, 4, 6)
    e_mine, e_eager = rel(out["out"], ref), rel(eager, ref)
    print(f"ControlNeXt rel err vs fp32 oracle: x2i_b200 {e_mine:.4f}, eager-bf16 reference path {e_eager:.4f}")
    assert e_mine < TOL
    assert rel(out["out"], d["out"]) < 1.5 * TOL  # the reference class's own output (fp32 weights before bf16 rounding)
    # a 256 x 256 hint (16 x 16 control tokens), batch 1, random timestep
    g = torch.Generator().manual_seed(3)
    hint2 = (torch.rand(1, 3, 256, 256, generator=g) * 2 - 1).to(torch.bfloat16)
    t2 = torch.tensor([504.0]).to(torch.bfloat16)
    with torch.no_grad():
        ref2 = net_o.float().cpu()(hint2.float(), t2.float())["out"]
        out2 = net.forward_tokens(hint2.cuda(), t2.cuda())
    assert out2.shape == (1, 256, 3072)
    assert rel(out2, ref2.flatten(2).transpose(1, 2)) < TOL


@gpu
def test_transformer_injection_matches_oracle(ops):
    """FluxTransformer2DModel(..., guided_hint=, control_nets=) at the real width (the control signal is 3072 wide)."""
    from oracle import flux_oracle as fo
    from x2i_b200.controlnext import ControlNeXtModel
    from x2i_b200.flux import FluxTransformer2DModel
    cfg = dict(patch_size=1, in_channels=64, num_layers=2, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
               joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    oracle = fo.FluxTransformer2DModel(**cfg).eval()
    fo.init_synthetic_(oracle, seed=61, std=0.02)
    nets_o = [_oracle_net(62, 0.05)]
    with torch.no_grad():
        for m in [oracle] + nets_o:
            for p_ in m.parameters():
                p_.copy_(p_.to(torch.bfloat16).float())
    model = FluxTransformer2DModel(**cfg).eval()
    model.load_state_dict(oracle.state_dict())
    model = model.to("cuda", torch.bfloat16)
    nets = torch.nn.ModuleList([ControlNeXtModel().eval() for _ in nets_o])
    for n, o in zip(nets, nets_o):
        n.load_state_dict(o.state_dict())
    nets = nets.to("cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(63)
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    B, hl, wl, S = 2, 4, 6, 8
    hint = bf(torch.rand(B, 3, 16 * hl, 16 * wl, generator=g) * 2 - 1)
    inp = dict(hidden_states=bf(torch.randn(B, hl * wl, 64, generator=g)), encoder_hidden_states=bf(torch.randn(B, S, 64, generator=g)),
               pooled_projections=bf(torch.randn(B, 32, generator=g)), timestep=torch.tensor([1.0, 0.5]),
               img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), txt_ids=torch.zeros(S, 3), guidance=torch.tensor([3.5, 3.5]))
    oin = dict(inp)
    for k in ("timestep", "guidance"):
        oin[k] = (inp[k].to(torch.bfloat16) * 1000).float() / 1000
    with torch.no_grad():
        ref = oracle(**oin, guided_hint=hint, control_nets=nets_o, return_dict=False)[0]
        ref0 = oracle(**oin, return_dict=False)[0]
        dev = {k: (v.to("cuda", torch.bfloat16) if k in ("hidden_states", "encoder_hidden_states", "pooled_projections") else v.cuda())
               for k, v in inp.items()}
        out = model(**dev, guided_hint=hint.to("cuda", torch.bfloat16), control_nets=nets, return_dict=False)[0]

        class Foreign(torch.nn.Module):  # any control net honouring the reference protocol goes through the generic path
            def __init__(self, inner):
                super().__init__()
                self.inner = inner

            def forward(self, sample, timestep):
                return self.inner(sample, timestep)

        out_f = model(**dev, guided_hint=hint.to("cuda", torch.bfloat16), control_nets=[Foreign(nets[0])], return_dict=False)[0]
    assert rel(ref0, ref) > 1e-2
    assert rel(out, ref) < TOL
    assert rel(out_f, out) < 5e-3


@gpu
def test_stacked_control_nets_equal_per_net_evaluation(ops):
    """ControlNeXtStack (one launch per layer for all nets, per-net weight sets) against the same nets evaluated one by one:
    same kernels, same per-image arithmetic -> bit-identical; and the transformer gives the same result with stacking on/off."""
    from x2i_b200.controlnext import ControlNeXtModel, ControlNeXtStack
    from x2i_b200.flux import FluxTransformer2DModel, init_synthetic_
    G, B = 3, 2
    nets = torch.nn.ModuleList([ControlNeXtModel().eval() for _ in range(G)]).to("cuda", torch.bfloat16)
    for i, n in enumerate(nets):
        init_synthetic_(n, seed=70 + i, std=0.05)
    g = torch.Generator(device="cuda").manual_seed(5)
    hint = (torch.rand(B, 3, 96, 64, device="cuda", generator=g) * 2 - 1).bfloat16()
    t = torch.tensor([700.0, 123.0], device="cuda")
    assert ControlNeXtStack.supported(nets)
    with torch.no_grad():
        mids = ControlNeXtStack(nets).mid_features(hint, t)
        assert mids.shape == (G, B, 12, 8, 256)
        for i, n in enumerate(nets):
            x0 = torch.randn(B, 6 * 4, 3072, device="cuda", generator=g).bfloat16()
            a, b = x0.clone(), x0.clone()
            n.forward_tokens(hint, t, add_to=a)
            n.finish_tokens(mids[i], add_to=b)
            assert torch.equal(a, b), f"net {i}"
            assert torch.equal(n.finish_tokens(mids[i]), n.forward_tokens(hint, t))
        # the hint-only part (embedding stack + first norm) is kept across the steps of a sampling run: a second timestep through the
        # cache equals a fresh evaluation, a modified hint or parameter invalidates it
        st = ControlNeXtStack(nets)
        t2 = torch.tensor([40.0, 910.0], device="cuda")
        st.mid_features(hint, t)
        cached = st.mid_features(hint, t2)
        assert st._hint is not None
        assert torch.equal(cached, ControlNeXtStack(nets).mid_features(hint, t2))
        hint2 = hint.clone()
        hint2[:, :, :8] = 0.5
        hint.copy_(hint2)                                                      # in-place change: same address, new version
        assert torch.equal(st.mid_features(hint, t2), ControlNeXtStack(nets).mid_features(hint2, t2))
        nets[1].embedding[3].weight.data.mul_(1.5)                            # .data bypasses the version counter: use a real update
        nets[1].embedding[3].weight.mul_(1.0)
        assert torch.equal(st.mid_features(hint, t2), ControlNeXtStack(nets).mid_features(hint, t2))
    cfg = dict(patch_size=1, in_channels=64, num_layers=3, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
               joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    model = FluxTransformer2DModel.synthetic(cfg, device="cuda", seed=9)
    from x2i_b200.pipeline import FluxPipeline
    kw = dict(hidden_states=torch.randn(B, 24, 64, device="cuda", generator=g).bfloat16(),
              encoder_hidden_states=torch.randn(B, 8, 64, device="cuda", generator=g).bfloat16(),
              pooled_projections=torch.randn(B, 32, device="cuda", generator=g).bfloat16(), timestep=torch.tensor([0.7, 0.2], device="cuda"),
              guidance=torch.tensor([3.5, 3.5], device="cuda"), txt_ids=torch.zeros(8, 3, device="cuda"),
              img_ids=FluxPipeline._prepare_latent_image_ids(B, 12, 8, "cuda", torch.float32), guided_hint=hint, control_nets=nets)
    with torch.no_grad():
        model.stack_control_nets = True
        y1 = model(**kw, return_dict=False)[0].clone()
        model.stack_control_nets = False
        y2 = model(**kw, return_dict=False)[0].clone()
    assert torch.equal(y1, y2)
    assert not ControlNeXtStack.supported(nets[:1])


@gpu
def test_controlnext_backward_matches_oracle_autograd(ops):
    """The trainable part of the LightControl trainer (train_lightcontrol.py:517-522): every parameter gradient of a ControlNeXtModel
    for a random upstream gradient, against torch autograd through the fp32 oracle on the same (bf16-rounded) weights.
    bf16 activations through a ~20-layer chain: per parameter tensor 3e-2 relative or 1.5x the deviation of the reference's own
    eager-bf16 autograd path (tiny bias gradients are noisy in both), and the same yardstick on the total."""
    from x2i_b200.controlnext import ControlNeXtModel
    net_o = _oracle_net(81, 0.05)
    with torch.no_grad():
        for p_ in net_o.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    net = ControlNeXtModel()
    net.load_state_dict(net_o.state_dict())
    net = net.to("cuda", torch.bfloat16).train()
    g = torch.Generator().manual_seed(82)
    hint = (torch.rand(2, 3, 64, 96, generator=g) * 2 - 1).to(torch.bfloat16)
    t = torch.tensor([700.0, 33.0])
    dout = torch.randn(2, 4 * 6, 3072, generator=g).to(torch.bfloat16)
    net_o = net_o.cuda()
    ref = net_o(hint.float().cuda(), t.cuda())["out"].flatten(2).transpose(1, 2)
    (ref * dout.float().cuda()).sum().backward()
    out = net.forward_tokens(hint.cuda(), t.cuda())
    assert out.requires_grad and out.shape == (2, 24, 3072)
    assert rel(out, ref.detach()) < TOL
    (out.float() * dout.float().cuda()).sum().backward()
    # yardstick: the reference's own bf16 path (oracle modules in bf16, torch eager autograd) against the same fp32 gradients
    import copy
    net_b = copy.deepcopy(net_o).to(torch.bfloat16)
    net_b.zero_grad()
    outb = net_b(hint.cuda(), t.cuda().to(torch.bfloat16))["out"].flatten(2).transpose(1, 2)
    (outb.float() * dout.float().cuda()).sum().backward()
    worst, num, den, worst_eager, num_e = 0.0, 0.0, 0.0, 0.0, 0.0
    for (n_, p_), (_, q_), (_, b_) in zip(net.named_parameters(), net_o.named_parameters(), net_b.named_parameters()):
        assert p_.grad is not None, n_
        e, eb = rel(p_.grad, q_.grad), rel(b_.grad, q_.grad)
        worst, worst_eager = max(worst, e), max(worst_eager, eb)
        num += float((p_.grad.float() - q_.grad).norm() ** 2)
        num_e += float((b_.grad.float() - q_.grad).norm() ** 2)
        den += float(q_.grad.norm() ** 2)
        # per-tensor numbers are recorded for every parameter; the bar is enforced on tensors big enough for a relative error to
        # be a statistic rather than noise (a 64-entry bias gradient is a sum of bf16-rounded terms in both paths), and on the
        # total over all parameters below
        if p_.numel() >= 4096:
            check(f"ControlNeXt backward: {n_}", e, eb)
        else:
            record(f"ControlNeXt backward: {n_} ({p_.numel()} elements, recorded only)", e, eb)
    tot, tot_e = (num / den) ** 0.5, (num_e / den) ** 0.5
    print(f"ControlNeXt backward: worst per-parameter rel err {worst:.4f} (eager bf16 {worst_eager:.4f}), total {tot:.4f} (eager bf16 {tot_e:.4f})")
    check("ControlNeXt backward: all parameter gradients", tot, tot_e)


@gpu
def test_lightcontrol_gradients_through_frozen_transformer(ops):
    """train_lightcontrol.py:732-775 in miniature: the frozen FLUX (real width, 2 double + 1 single block) with 2 trainable control nets;
    d loss / d (every control-net parameter) through the hand-written transformer backward + the ControlNeXt backward, against torch
    autograd through the fp32 oracle pipeline (yardstick: the same pipeline in eager bf16)."""
    import copy
    from oracle import flux_oracle as fo
    from x2i_b200.controlnext import ControlNeXtModel
    from x2i_b200.flux import FluxTransformer2DModel
    cfg = dict(patch_size=1, in_channels=64, num_layers=2, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
               joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    oracle = fo.FluxTransformer2DModel(**cfg).eval()
    fo.init_synthetic_(oracle, seed=91, std=0.02)
    nets_o = [_oracle_net(92, 0.05), _oracle_net(93, 0.05)]
    with torch.no_grad():
        for m in [oracle] + nets_o:
            for p_ in m.parameters():
                p_.copy_(p_.to(torch.bfloat16).float())
    model = FluxTransformer2DModel(**cfg).eval()
    model.load_state_dict(oracle.state_dict())
    model = model.to("cuda", torch.bfloat16).requires_grad_(False)
    nets = torch.nn.ModuleList([ControlNeXtModel() for _ in nets_o])
    for n, o in zip(nets, nets_o):
        n.load_state_dict(o.state_dict())
    nets = nets.to("cuda", torch.bfloat16).train()
    g = torch.Generator().manual_seed(94)
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    B, hl, wl, S = 2, 4, 6, 8
    hint = bf(torch.rand(B, 3, 16 * hl, 16 * wl, generator=g) * 2 - 1)
    inp = dict(hidden_states=bf(torch.randn(B, hl * wl, 64, generator=g)), encoder_hidden_states=bf(torch.randn(B, S, 64, generator=g)),
               pooled_projections=bf(torch.randn(B, 32, generator=g)), timestep=torch.tensor([1.0, 0.5]),
               img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), txt_ids=torch.zeros(S, 3), guidance=torch.tensor([3.5, 3.5]))
    target = bf(torch.randn(B, hl * wl, 64, generator=g))
    oin = dict(inp)
    for k in ("timestep", "guidance"):
        oin[k] = (inp[k].to(torch.bfloat16) * 1000).float() / 1000

    def oracle_loss(tr, cn, dtype):
        kw = {k: (v.cuda().to(dtype) if v.is_floating_point() else v.cuda()) for k, v in oin.items()}
        out = tr(**kw, guided_hint=hint.cuda().to(dtype), control_nets=cn, return_dict=False)[0]
        return ((out.float() - target.cuda()) ** 2).mean()

    oracle, nets_o = oracle.cuda().requires_grad_(False), [n.cuda() for n in nets_o]
    lo = oracle_loss(oracle, nets_o, torch.float32)
    lo.backward()
    ob, nb = copy.deepcopy(oracle).to(torch.bfloat16), [copy.deepcopy(n).to(torch.bfloat16) for n in nets_o]
    for n in nb:
        n.zero_grad()
    lb = oracle_loss(ob, nb, torch.bfloat16)
    lb.backward()
    dev = {k: (v.to("cuda", torch.bfloat16) if k in ("hidden_states", "encoder_hidden_states", "pooled_projections") else v.cuda())
           for k, v in inp.items()}
    out = model(**dev, guided_hint=hint.to("cuda", torch.bfloat16), control_nets=nets, return_dict=False)[0]
    assert out.requires_grad
    loss = ((out.float() - target.cuda()) ** 2).mean()
    loss.backward()
    assert abs(float(loss) - float(lo)) / float(lo) < 1e-2
    num = den = num_e = 0.0
    for net, no, nbb in zip(nets, nets_o, nb):
        for (n_, p_), (_, q_), (_, b_) in zip(net.named_parameters(), no.named_parameters(), nbb.named_parameters()):
            assert p_.grad is not None and torch.isfinite(p_.grad.float()).all(), n_
            num += float((p_.grad.float() - q_.grad).norm() ** 2)
            num_e += float((b_.grad.float() - q_.grad).norm() ** 2)
            den += float(q_.grad.norm() ** 2)
    tot, tot_e = (num / den) ** 0.5, (num_e / den) ** 0.5
    print(f"LightControl gradients: total rel err {tot:.4f} (eager bf16 {tot_e:.4f}); loss {float(loss):.5f} vs {float(lo):.5f}")
    check("LightControl gradients through the frozen real-width transformer", tot, tot_e)
    assert all(p.grad is None for p in model.parameters())  # the transformer stays frozen
    # the control nets run round-robin on side streams by default (forward and, through autograd's stream tracking, backward): same
    # kernels on the same data -> the single-stream run gives bit-identical gradients
    multi = [p_.grad.clone() for net in nets for p_ in net.parameters()]
    assert model.control_net_streams > 1
    for net in nets:
        net.zero_grad()
    model.control_net_streams = 1
    try:
        out1 = model(**dev, guided_hint=hint.to("cuda", torch.bfloat16), control_nets=nets, return_dict=False)[0]
        ((out1.float() - target.cuda()) ** 2).mean().backward()
    finally:
        model.control_net_streams = 8
    assert torch.equal(out1, out)
    for a_, p_ in zip(multi, [p_ for net in nets for p_ in net.parameters()]):
        assert torch.equal(a_, p_.grad)


@gpu
def test_lightcontrol_train_step_end_to_end(ops):
    """x2i_b200.train_lightcontrol.lightcontrol_step (train_lightcontrol.py:672-775): VAE encode -> flow-matching inputs -> frozen FLUX
    + trainable control nets -> MSE -> backward -> clip -> AdamW.  Repeating one batch with fixed noise / timesteps must reduce the
    loss; only the control nets change."""
    from x2i_b200 import train_lightcontrol as tl, vae as xv
    from x2i_b200.controlnext import ControlNeXtModel
    from x2i_b200.flux import FluxTransformer2DModel, init_synthetic_
    cfg = dict(patch_size=1, in_channels=64, num_layers=2, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
               joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    model = FluxTransformer2DModel.synthetic(cfg, device="cuda", seed=101).requires_grad_(False)
    vae = init_synthetic_(xv.AutoencoderKL(block_out_channels=(64, 64, 128, 128), norm_num_groups=16).to("cuda", torch.bfloat16).eval(),
                          seed=102, std=0.05).requires_grad_(False)
    nets = torch.nn.ModuleList([ControlNeXtModel() for _ in range(2)]).to("cuda", torch.bfloat16).train()
    for i, n in enumerate(nets):
        init_synthetic_(n, seed=103 + i, std=0.05)
    before = [p.detach().clone() for p in nets.parameters()]
    frozen = [p.detach().clone() for p in list(model.parameters())[:4] + list(vae.parameters())[:4]]
    g = torch.Generator(device="cuda").manual_seed(7)
    B = 2
    batch = dict(pixel_values=(torch.rand(B, 3, 64, 96, device="cuda", generator=g) * 2 - 1).bfloat16(),
                 prompt_embeds=torch.randn(B, 8, 64, device="cuda", generator=g).bfloat16(),
                 pooled_prompt_embeds=torch.randn(B, 32, device="cuda", generator=g).bfloat16())
    opt = torch.optim.AdamW(nets.parameters(), lr=1e-4, weight_decay=0.0)  # Adam moves every weight by ~lr per step (std 0.05)
    losses = []
    for _ in range(8):
        gen = torch.Generator(device="cuda").manual_seed(11)  # same latent sample, noise and timesteps every step
        losses.append(float(tl.lightcontrol_step(nets, model, vae, batch, optimizer=opt, generator=gen,
                                                 sigmas=torch.tensor([0.8, 0.3]))))
    print("LightControl train losses:", [round(l, 4) for l in losses])
    assert all(l == l and l < 1e4 for l in losses) and losses[-1] < losses[0]
    assert any(not torch.equal(a, p.detach()) for a, p in zip(before, nets.parameters()))
    assert all(torch.equal(a, p.detach()) for a, p in zip(frozen, list(model.parameters())[:4] + list(vae.parameters())[:4]))
    assert all(p.grad is None for p in model.parameters()) and all(p.grad is None for p in vae.parameters())
    packed, ts, target, h, w = tl.flow_matching_inputs(vae, batch["pixel_values"], generator=torch.Generator(device="cuda").manual_seed(11))
    assert packed.shape == (B, 24, 64) and target.shape == (B, 16, 8, 12) and (h, w) == (8, 12) and float(ts.min()) >= 0 and float(ts.max()) <= 1000
