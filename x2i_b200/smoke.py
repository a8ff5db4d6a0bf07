"""One small invocation of the hot path on cuda:0, checked against the CPU oracle (called by __graft_entry__.smoke)."""
import os
import sys

import torch


def tiny_config(guidance=False):
    return dict(patch_size=1, in_channels=64, num_layers=2, num_single_layers=3, attention_head_dim=128,
                num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=guidance,
                axes_dims_rope=(16, 56, 56))


def make_pair(cfg, seed=0, device="cuda"):
    """(product model on `device` in bf16, oracle model on CPU in fp32) sharing the same bf16-rounded weights."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import flux_oracle as fo  # checker only
    from x2i_b200.flux import FluxTransformer2DModel
    oracle = fo.FluxTransformer2DModel(**cfg).eval()
    fo.init_synthetic_(oracle, seed=seed, std=0.05)
    with torch.no_grad():
        for p in oracle.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    model = FluxTransformer2DModel(**cfg).eval()
    model.load_state_dict(oracle.state_dict())
    model = model.to(device, torch.bfloat16)
    return model, oracle


def make_inputs(cfg, B=2, hl=8, wl=8, S=24, seed=1):
    from oracle import flux_oracle as fo
    g = torch.Generator().manual_seed(seed)
    bf = lambda t: t.to(torch.bfloat16).float()  # noqa: E731
    inp = dict(hidden_states=bf(torch.randn(B, hl * wl, cfg["in_channels"], generator=g)),
               encoder_hidden_states=bf(torch.randn(B, S, cfg["joint_attention_dim"], generator=g)),
               pooled_projections=bf(torch.randn(B, cfg["pooled_projection_dim"], generator=g)),
               timestep=torch.tensor([1.0, 0.75, 0.5, 0.25][:B] if B <= 4 else [0.5] * B),
               img_ids=fo.prepare_latent_image_ids(2 * hl, 2 * wl), txt_ids=torch.zeros(S, 3))
    if cfg["guidance_embeds"]:
        inp["guidance"] = torch.full((B,), 3.5)
    return inp


def oracle_inputs(inp):
    """Inputs for the fp32 oracle that reproduce the reference's bf16 quirk: it forms `timestep.to(bf16) * 1000` and
    `guidance.to(bf16) * 1000` IN bf16 (lightcontrol_flux.py:447-449), e.g. 0.75 -> 752 and 3.5 -> 3504.  The product
    path does the same; a pure-fp32 oracle must be told the rounded values or the time embedding differs by ~25%."""
    out = dict(inp)
    for k in ("timestep", "guidance"):
        if k in out and out[k] is not None:
            out[k] = (out[k].to(torch.bfloat16) * 1000).float() / 1000
    return out


def to_device(inp, device="cuda"):
    out = {}
    for k, v in inp.items():
        if k in ("hidden_states", "encoder_hidden_states", "pooled_projections"):
            out[k] = v.to(device, torch.bfloat16)
        else:
            out[k] = v.to(device)
    return out


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def run_smoke():
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    torch.cuda.set_device(0)
    cfg = tiny_config(guidance=True)
    model, oracle = make_pair(cfg)
    inp = make_inputs(cfg)
    with torch.no_grad():
        ref = oracle(**oracle_inputs(inp), return_dict=False)[0]
        out = model(**to_device(inp), return_dict=False)[0]
    torch.cuda.synchronize()
    err = rel(out, ref)
    from x2i_b200 import _lib
    print(f"[smoke] FLUX MMDiT step (2 double + 3 single blocks, D=256, L=24+64) rel err vs fp32 oracle = {err:.4f}; "
          f"kernel launches through libx2i_b200.so = {_lib.launch_count()}")
    assert err < 2e-2, f"smoke: rel err {err}"
