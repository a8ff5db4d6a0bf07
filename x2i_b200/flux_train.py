"""Differentiable path through the FLUX MMDiT for attention-distillation training.

The reference's student pass (``train/train_qwenvl.py:576-632``) is: projector -> frozen ``FluxTransformer2DModel`` with
forward hooks on every ``blk.attn`` -> KD loss over the hooked tensors -> ``loss.backward()`` -> projector update.  The
FLUX weights are frozen (``:417-429``), so the backward through the 57 blocks needs only ACTIVATION gradients; the only
weight gradients of the step belong to the projector.

Here the whole transformer is ONE ``torch.autograd.Function``: its forward runs the same sm_100a kernels as inference in
a mode that keeps what the backward needs (block inputs, pre-norm q|k, post-RoPE q/k/v, attention output + row
log-sum-exp, pre-GELU MLP activations, un-gated branch outputs), its outputs are the model output PLUS every hooked
tensor (so the reference's hooks receive tensors that carry autograd history, B3 contract), and its backward walks the
blocks in reverse with the hand-written backward kernels (dgrad GEMMs on the weights as stored, the tcgen05 attention
backward, fused row-wise kernels), injecting each hook tensor's incoming gradient at the block it belongs to.
Gradients are returned for ``encoder_hidden_states`` (-> projector sequence output) and ``pooled_projections``
(-> projector pooled output, through temb and all 77 AdaLN modulation linears), and for ``hidden_states``.

PyTorch is plumbing (buffers + the autograd graph edge); there is no eager fallback.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import X2IError

BF16 = torch.bfloat16
F32 = torch.float32


def _e(*shape, device, dtype=BF16):
    return torch.empty(*shape, device=device, dtype=dtype)


# ------------------------------------------------------------------------------------------------ forward (saving)
def _double_fwd(blk, x0, c0, mod, rope, B, L_img, S):
    """x0 [B*L_img, D], c0 [B*S, D] -> (x2, c2, saved).  lightcontrol_flux.py:159-204."""
    D = blk.dim
    dev = x0.device
    at = blk.attn
    H = at.heads
    L = S + L_img
    mi = [mod[:, i * D:(i + 1) * D] for i in range(6)]
    mc = [mod[:, (6 + i) * D:(7 + i) * D] for i in range(6)]
    nx = ops.ln_modulate(x0, mi[1], mi[0], L_img)
    nc = ops.ln_modulate(c0, mc[1], mc[0], S)
    q = _e(B, H, L, 128, device=dev); k = torch.empty_like(q); v = torch.empty_like(q)
    qk_pre_i, qk_pre_t = _e(B * L_img, 2 * D, device=dev), _e(B * S, 2 * D, device=dev)
    ops.gemm_grouped(
        ops.desc_qkv_rope_save(nx, at._w_qkv, at._b_qkv, at.norm_q.weight, at.norm_k.weight, rope, q, k, v, H, L_img, S, qk_pre_i,
                               at.norm_q.eps),
        ops.desc_qkv_rope_save(nc, at._w_add_qkv, at._b_add_qkv, at.norm_added_q.weight, at.norm_added_k.weight, rope, q, k, v, H,
                               S, 0, qk_pre_t, at.norm_added_q.eps))
    a_txt, a_img, lse = ops.attention_lse(q, k, v, split=S)
    x1, c1, ya_i, ya_t = torch.empty_like(x0), torch.empty_like(c0), torch.empty_like(x0), torch.empty_like(c0)
    wo, wa = at.to_out[0], at.to_add_out
    ops.gemm_grouped(
        ops.desc_gate_residual(a_img.view(B * L_img, D), wo.weight, wo.bias, mi[2], x0, L_img, aux=ya_i, out=x1),
        ops.desc_gate_residual(a_txt.view(B * S, D), wa.weight, wa.bias, mc[2], c0, S, aux=ya_t, out=c1))
    nx2 = ops.ln_modulate(x1, mi[4], mi[3], L_img, out=nx)
    nc2 = ops.ln_modulate(c1, mc[4], mc[3], S, out=nc)
    f0, f2, g0, g2 = blk.ff.net[0].proj, blk.ff.net[2], blk.ff_context.net[0].proj, blk.ff_context.net[2]
    F = f0.weight.shape[0]
    hpre_i, hpre_t = _e(B * L_img, F, device=dev), _e(B * S, F, device=dev)
    hact_i, hact_t = torch.empty_like(hpre_i), torch.empty_like(hpre_t)
    ops.gemm_grouped(ops.desc_linear_act_save(nx2, f0.weight, f0.bias, hpre_i, hact_i, act=1),
                     ops.desc_linear_act_save(nc2, g0.weight, g0.bias, hpre_t, hact_t, act=1))
    x2, c2, yf_i, yf_t = torch.empty_like(x0), torch.empty_like(c0), torch.empty_like(x0), torch.empty_like(c0)
    ops.gemm_grouped(ops.desc_gate_residual(hact_i, f2.weight, f2.bias, mi[5], x1, L_img, aux=yf_i, out=x2),
                     ops.desc_gate_residual(hact_t, g2.weight, g2.bias, mc[5], c1, S, aux=yf_t, out=c2))
    saved = dict(x0=x0, c0=c0, x1=x1, c1=c1, qk_pre_i=qk_pre_i, qk_pre_t=qk_pre_t, q=q, k=k, v=v, a_img=a_img, a_txt=a_txt,
                 lse=lse, ya_i=ya_i, ya_t=ya_t, hpre_i=hpre_i, hpre_t=hpre_t, yf_i=yf_i, yf_t=yf_t)
    return x2, c2, saved


def _single_fwd(blk, h0, mod, rope, B, L, cat_buf):
    """h0 [B*L, D] -> (h1, saved).  lightcontrol_flux.py:82-104."""
    D, F = blk.dim, blk.mlp_hidden_dim
    dev = h0.device
    at = blk.attn
    H = at.heads
    shift, scale, gate = (mod[:, i * D:(i + 1) * D] for i in range(3))
    n = ops.ln_modulate(h0, scale, shift, L)
    q = _e(B, H, L, 128, device=dev); k = torch.empty_like(q); v = torch.empty_like(q)
    qk_pre, mlp_pre = _e(B * L, 2 * D, device=dev), _e(B * L, F, device=dev)
    ops.gemm_grouped(ops.desc_qkv_rope_save(n, at._w_qkv_mlp, at._b_qkv_mlp, at.norm_q.weight, at.norm_k.weight, rope, q, k, v, H, L,
                                            0, qk_pre, at.norm_q.eps, mlp=cat_buf[:, D:], mlp_pre=mlp_pre))
    _, _, lse = ops.attention_lse(q, k, v, split=0, out1=cat_buf.view(B, L, D + F)[:, :, :D])
    a = cat_buf.view(B, L, D + F)[:, :, :D].contiguous()  # the hooked tensor (train_qwenvl.py:214) and O of the backward
    h1, y = torch.empty_like(h0), torch.empty_like(h0)
    ops.linear_gate_residual(cat_buf, blk.proj_out.weight, blk.proj_out.bias, gate, h0, L, out=h1, aux=y)
    return h1, dict(h0=h0, qk_pre=qk_pre, mlp_pre=mlp_pre, q=q, k=k, v=v, a=a, lse=lse, y=y)


def forward_save(model, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids, guidance, controls=()):
    """Training-mode forward.  Returns (out [B, L_img, C], hooks, tape).  controls: the LightControl control tokens [B, L_img, D],
    added to the image stream after double block i (lightcontrol_flux.py:504-507)."""
    B, L_img, _ = hidden_states.shape
    S = encoder_hidden_states.shape[1]
    D = model.inner_dim
    L = S + L_img
    dev = hidden_states.device
    _, rope = model._rope(txt_ids, img_ids)
    x = ops.linear(hidden_states.to(BF16).contiguous(), model.x_embedder.weight, model.x_embedder.bias).view(B * L_img, D)
    t1000 = timestep.to(BF16) * 1000
    tte = model.time_text_embed
    temb = tte._mlp(tte.timestep_embedder, ops.timestep_sinusoid(t1000.float().contiguous()))
    if guidance is not None:
        tte._mlp(tte.guidance_embedder, ops.timestep_sinusoid((guidance.to(BF16) * 1000).float().contiguous()), out=temb, accumulate=True)
    pooled_b = pooled.to(BF16).contiguous()
    z1 = ops.skinny_linear(pooled_b, tte.text_embedder.linear_1.weight, tte.text_embedder.linear_1.bias)
    ops.skinny_linear(z1, tte.text_embedder.linear_2.weight, tte.text_embedder.linear_2.bias, act_in=1, out=temb, accumulate=True)
    c = ops.linear(encoder_hidden_states.to(BF16).contiguous(), model.context_embedder.weight, model.context_embedder.bias).view(B * S, D)
    mod = ops.skinny_linear(temb, model._w_mod, model._b_mod, act_in=1)

    # Gradient checkpointing (lightcontrol_flux.py:475-494,:513-531, enabled by train_lightcontrol.py:666): keep only each
    # block's inputs (28-31 MB per sample instead of ~0.35 GB) and re-run the block's forward kernels in the backward.
    ckpt = bool(getattr(model, "gradient_checkpointing", False))
    tape = dict(B=B, S=S, L_img=L_img, rope=rope, temb=temb, z1=z1, mod=mod, double=[], single=[],
                enc_dim=encoder_hidden_states.shape[2], in_dim=hidden_states.shape[2])
    hooks_img, hooks_txt, hooks_single = [], [], []
    off = 0
    for blk in model.transformer_blocks:
        x0, c0 = x, c
        x, c, sv = _double_fwd(blk, x, c, mod[:, off:off + 12 * D], rope, B, L_img, S)
        tape["double"].append(dict(x0=x0, c0=c0, recompute=True) if ckpt else sv)
        hooks_img.append(sv["ya_i"].view(B, L_img, D))
        hooks_txt.append(sv["ya_t"].view(B, S, D))
        off += 12 * D
        i_blk = len(tape["double"]) - 1
        if i_blk < len(controls):  # hidden_states += control['out'] * scale (scale == 1.0); x is block i+1's saved input
            ops.euler_step_(x, controls[i_blk].to(BF16).contiguous().view(B * L_img, D), 1.0)
    h = _e(B, L, D, device=dev)
    h[:, :S].copy_(c.view(B, S, D))
    h[:, S:].copy_(x.view(B, L_img, D))
    h = h.view(B * L, D)
    if len(model.single_transformer_blocks):
        F = model.single_transformer_blocks[0].mlp_hidden_dim
        cat_buf = _e(B * L, D + F, device=dev)
    for blk in model.single_transformer_blocks:
        h0 = h
        h, sv = _single_fwd(blk, h, mod[:, off:off + 3 * D], rope, B, L, cat_buf)
        tape["single"].append(dict(h0=h0, recompute=True) if ckpt else sv)
        hooks_single.append(sv["a"])
        off += 3 * D
    tape["h_final"] = h
    tape["off_out"] = off
    n = ops.ln_modulate(h, mod[:, off:off + D], mod[:, off + D:off + 2 * D], L)
    out = ops.linear(n, model.proj_out.weight, model.proj_out.bias).view(B, L, -1)[:, S:].contiguous()
    return out, (hooks_img, hooks_txt, hooks_single), tape


# ------------------------------------------------------------------------------------------------ backward
def _g2(t, rows, D):
    """Incoming hook gradient as a contiguous bf16 [rows, D] matrix (None stays None)."""
    if t is None:
        return None
    return t.to(BF16).contiguous().view(rows, D)


def _double_bwd(blk, sv, dx2, dc2, gkd_i, gkd_t, mod, dmod, rope, B, L_img, S):
    """Gradients w.r.t. the block inputs (dx0, dc0); the modulation gradients are written into dmod (fp32 view)."""
    D = blk.dim
    at = blk.attn
    H = at.heads
    L = S + L_img
    mi = [mod[:, i * D:(i + 1) * D] for i in range(6)]
    mc = [mod[:, (6 + i) * D:(7 + i) * D] for i in range(6)]
    di = [dmod[:, i * D:(i + 1) * D] for i in range(6)]     # d(shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp)
    dc = [dmod[:, (6 + i) * D:(7 + i) * D] for i in range(6)]
    f0, f2, g0, g2 = blk.ff.net[0].proj, blk.ff.net[2], blk.ff_context.net[0].proj, blk.ff_context.net[2]
    dev = dx2.device

    def mlp_branch(dxo, m, dm, rows, yf, hpre, w0, w2, x1):
        ops.colsum(dxo, B, rows, b=yf, out1=dm[5])                        # dgate_mlp = sum dx' * y_ff
        dyf = ops.gate_bwd(dxo, m[5], rows)
        dh = ops.linear_dgrad(dyf, w2.weight, pre=hpre, n_split=0, dact=1)  # through Linear2 and GELU(tanh)
        dn2 = ops.linear_dgrad(dh, w0.weight, out=dyf)
        stats = _e(dxo.shape[0], 2, device=dev, dtype=F32)
        dx1 = ops.ln_modulate_bwd(dn2, x1, m[4], rows, dres=dxo, stats=stats)
        ops.colsum(dn2, B, rows, out0=dm[3], b=x1, out1=dm[4], stats=stats)
        return dx1

    dx1 = mlp_branch(dx2, mi, di, L_img, sv["yf_i"], sv["hpre_i"], f0, f2, sv["x1"])
    dc1 = mlp_branch(dc2, mc, dc, S, sv["yf_t"], sv["hpre_t"], g0, g2, sv["c1"])

    def out_proj(dx1_, m, dm, rows, ya, gkd, w):
        ops.colsum(dx1_, B, rows, b=ya, out1=dm[2])                        # dgate_msa
        dya = ops.gate_bwd(dx1_, m[2], rows, addend=gkd)                   # + KD gradient at the hooked tensor
        return ops.linear_dgrad(dya, w.weight)

    da_img = out_proj(dx1, mi, di, L_img, sv["ya_i"], gkd_i, at.to_out[0])
    da_txt = out_proj(dc1, mc, dc, S, sv["ya_t"], gkd_t, at.to_add_out)
    do_hm, delta = ops.attention_bwd_prep(da_txt.view(B, S, D), da_img.view(B, L_img, D), sv["a_txt"], sv["a_img"], B, H, L, S)
    dq, dk, dv = ops.attention_bwd(sv["q"], sv["k"], sv["v"], do_hm, sv["lse"], delta)
    dqkv_i, dqkv_t = _e(B * L_img, 3 * D, device=dev), _e(B * S, 3 * D, device=dev)
    ops.qk_norm_rope_bwd(dq, dk, dv, sv["qk_pre_i"], at.norm_q.weight, at.norm_k.weight, rope, dqkv_i, L_img, S, at.norm_q.eps)
    ops.qk_norm_rope_bwd(dq, dk, dv, sv["qk_pre_t"], at.norm_added_q.weight, at.norm_added_k.weight, rope, dqkv_t, S, 0,
                         at.norm_added_q.eps)

    def ln1(dqkv, w, x0_, m, dm, rows, dres):
        dn1 = ops.linear_dgrad(dqkv, w)
        stats = _e(dn1.shape[0], 2, device=dev, dtype=F32)
        dx0 = ops.ln_modulate_bwd(dn1, x0_, m[1], rows, dres=dres, stats=stats)
        ops.colsum(dn1, B, rows, out0=dm[0], b=x0_, out1=dm[1], stats=stats)
        return dx0

    dx0 = ln1(dqkv_i, at._w_qkv, sv["x0"], mi, di, L_img, dx1)
    dc0 = ln1(dqkv_t, at._w_add_qkv, sv["c0"], mc, dc, S, dc1)
    return dx0, dc0


def _single_bwd(blk, sv, dh1, gkd, mod, dmod, rope, B, L, dbig):
    D, F = blk.dim, blk.mlp_hidden_dim
    at = blk.attn
    H = at.heads
    dev = dh1.device
    scale, gate = mod[:, D:2 * D], mod[:, 2 * D:3 * D]
    ops.colsum(dh1, B, L, b=sv["y"], out1=dmod[:, 2 * D:3 * D])                                 # dgate
    dy = ops.gate_bwd(dh1, gate, L)
    # through proj_out on [attn | gelu(mlp)]: columns [2D,3D) of dbig receive d(attn) for now, [3D, 3D+F) d(mlp_pre)
    ops.linear_dgrad(dy, blk.proj_out.weight, pre=sv["mlp_pre"], n_split=D, dact=1, out=dbig[:, 2 * D:])
    do_hm, delta = ops.attention_bwd_prep(None, dbig[:, 2 * D:3 * D], None, sv["a"], B, H, L, 0, add1=gkd)
    dq, dk, dv = ops.attention_bwd(sv["q"], sv["k"], sv["v"], do_hm, sv["lse"], delta)
    ops.qk_norm_rope_bwd(dq, dk, dv, sv["qk_pre"], at.norm_q.weight, at.norm_k.weight, rope, dbig, L, 0, at.norm_q.eps)
    dn = ops.linear_dgrad(dbig, at._w_qkv_mlp, out=dy)
    stats = _e(B * L, 2, device=dev, dtype=F32)
    dh0 = ops.ln_modulate_bwd(dn, sv["h0"], scale, L, dres=dh1, stats=stats)
    ops.colsum(dn, B, L, out0=dmod[:, 0:D], b=sv["h0"], out1=dmod[:, D:2 * D], stats=stats)
    return dh0


def backward(model, tape, dout, dhooks_img, dhooks_txt, dhooks_single, need_hidden_grad=False, n_controls=0):
    """Returns (d hidden_states | None, d encoder_hidden_states [B,S,E] bf16, d pooled_projections [B,P] bf16, d controls list):
    the gradient of control i is the image-stream gradient at the output of double block i."""
    B, S, L_img = tape["B"], tape["S"], tape["L_img"]
    D = model.inner_dim
    L = S + L_img
    mod, rope = tape["mod"], tape["rope"]
    dev = mod.device
    dmod = torch.zeros(B, mod.shape[1], device=dev, dtype=F32)
    off = tape["off_out"]
    if dout is not None:
        C = dout.shape[-1]
        dfull = torch.zeros(B, L, C, device=dev, dtype=BF16)
        dfull[:, S:].copy_(dout)
        dn = ops.linear_dgrad(dfull.view(B * L, C), model.proj_out.weight)
        stats = _e(B * L, 2, device=dev, dtype=F32)
        dh = ops.ln_modulate_bwd(dn, tape["h_final"], mod[:, off:off + D], L, stats=stats)
        ops.colsum(dn, B, L, out0=dmod[:, off + D:off + 2 * D], b=tape["h_final"], out1=dmod[:, off:off + D], stats=stats)
    else:
        dh = torch.zeros(B * L, D, device=dev, dtype=BF16)
    if len(model.single_transformer_blocks):
        F = model.single_transformer_blocks[0].mlp_hidden_dim
        dbig = _e(B * L, 3 * D + F, device=dev)
    cat_buf = None
    for i in range(len(model.single_transformer_blocks) - 1, -1, -1):
        off -= 3 * D
        sv = tape["single"][i]
        if sv.get("recompute"):  # checkpointed block: rebuild its saved activations from the block input
            if cat_buf is None:
                cat_buf = _e(B * L, D + F, device=dev)
            _, sv = _single_fwd(model.single_transformer_blocks[i], sv["h0"], mod[:, off:off + 3 * D], rope, B, L, cat_buf)
        dh = _single_bwd(model.single_transformer_blocks[i], sv, dh, _g2(dhooks_single[i], B * L, D),
                         mod[:, off:off + 3 * D], dmod[:, off:off + 3 * D], rope, B, L, dbig)
        sv = None
        tape["single"][i] = None  # release this block's activations
    dh3 = dh.view(B, L, D)
    dc = dh3[:, :S].contiguous().view(B * S, D)
    dx = dh3[:, S:].contiguous().view(B * L_img, D)
    dctrl = [None] * n_controls
    for i in range(len(model.transformer_blocks) - 1, -1, -1):
        off -= 12 * D
        if i < n_controls:
            dctrl[i] = dx.view(B, L_img, D).clone()
        sv = tape["double"][i]
        if sv.get("recompute"):
            _, _, sv = _double_fwd(model.transformer_blocks[i], sv["x0"], sv["c0"], mod[:, off:off + 12 * D], rope, B, L_img, S)
        dx, dc = _double_bwd(model.transformer_blocks[i], sv, dx, dc, _g2(dhooks_img[i], B * L_img, D),
                             _g2(dhooks_txt[i], B * S, D), mod[:, off:off + 12 * D], dmod[:, off:off + 12 * D], rope, B, L_img, S)
        sv = None
        tape["double"][i] = None
    d_enc = ops.linear_dgrad(dc, model.context_embedder.weight).view(B, S, tape["enc_dim"])
    d_hidden = ops.linear_dgrad(dx, model.x_embedder.weight).view(B, L_img, tape["in_dim"]) if need_hidden_grad else None
    # modulation -> temb -> pooled text projection (PixArtAlphaTextProjection: Linear, SiLU, Linear)
    tte = model.time_text_embed.text_embedder
    dtemb = ops.skinny_linear_t(dmod, model._w_mod, pre=tape["temb"], dact=1)
    dz1 = ops.skinny_linear_t(dtemb, tte.linear_2.weight, pre=tape["z1"], dact=1)
    d_pooled = ops.f32_to_bf16(ops.skinny_linear_t(dz1, tte.linear_1.weight))
    return d_hidden, d_enc, d_pooled, dctrl


class FluxTrainFn(torch.autograd.Function):
    """(hidden_states, encoder_hidden_states, pooled_projections) -> (out, *hook tensors)."""

    @staticmethod
    def forward(ctx, model, n_hooks, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids, guidance, *controls):
        out, (hi, ht, hs), tape = forward_save(model, hidden_states.detach(), encoder_hidden_states.detach(), pooled.detach(),
                                               timestep, img_ids, txt_ids, guidance, tuple(c.detach() for c in controls))
        ctx.ctrl_dtypes = tuple(c.dtype for c in controls)
        ctx.set_materialize_grads(False)  # hook tensors nobody used arrive as None, not as 28 MB of zeros
        ctx.model, ctx.tape = model, tape
        ctx.n = (len(hi), len(ht), len(hs))
        ctx.dtypes = (hidden_states.dtype, encoder_hidden_states.dtype, pooled.dtype)
        ctx.need_hidden = hidden_states.requires_grad
        return (out, *hi, *ht, *hs)

    @staticmethod
    def backward(ctx, dout, *dh):
        if ctx.tape is None:
            raise X2IError("FluxTrainFn: backward called twice (activations are released during the first backward)")
        ni, nt, ns = ctx.n
        d_hidden, d_enc, d_pooled, dctrl = backward(ctx.model, ctx.tape, dout, dh[:ni], dh[ni:ni + nt], dh[ni + nt:], ctx.need_hidden,
                                                    len(ctx.ctrl_dtypes))
        ctx.tape = None
        t0, t1, t2 = ctx.dtypes
        return (None, None, d_hidden.to(t0) if d_hidden is not None else None, d_enc.to(t1), d_pooled.to(t2), None, None, None, None,
                *[g.to(dt) for g, dt in zip(dctrl, ctx.ctrl_dtypes)])
