"""MLLM prefill with all-layer hidden-state capture (infer/inference_qwenvl.py:121-132,:176-179;
train/train_qwenvl.py:773-775).  Oracle = the in-image ``transformers`` Qwen2.5-VL text model driven like the reference drives it
(oracle/mllm_oracle.py).  CPU: host logic and naming; -m gpu: the new kernels and the drop-in against the oracle."""
import pytest
import torch

from parity import check, record

TINY = dict(vocab_size=1000, hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2, num_key_value_heads=1,
            rms_norm_eps=1e-6, rope_theta=1000000.0)
gpu = pytest.mark.gpu


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


# ------------------------------------------------------------------------------------------------ CPU
def test_parameter_names_match_transformers_and_checkpoint_prefixes_load():
    from oracle import mllm_oracle as mo
    from x2i_b200 import mllm
    o = mo.build(TINY)
    with torch.device("meta"):
        m = mllm.Qwen2_5_VLTextPrefill(**TINY)
    want = {k: tuple(v.shape) for k, v in o.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    m = mllm.Qwen2_5_VLTextPrefill(**TINY)
    for prefix in ("model.language_model.", "model.", ""):  # transformers >= 4.52, 4.49 (the reference's pin), bare text model
        sd = {prefix + k: v for k, v in o.state_dict().items()}
        sd["visual.blocks.0.attn.qkv.weight"] = torch.zeros(1)
        sd["lm_head.weight"] = torch.zeros(1)
        m.load_hf_state_dict(sd)
        assert torch.equal(m.layers[2].mlp.down_proj.weight, o.layers[2].mlp.down_proj.weight)
    assert mllm.QWEN2_5_VL_3B["num_hidden_layers"] + 1 == 37 and mllm.QWEN2_5_VL_7B["num_hidden_layers"] + 1 == 29  # C of the projectors


def test_text_positions_and_left_padding_rule():
    from oracle import mllm_oracle as mo
    from x2i_b200 import mllm
    from x2i_b200._lib import X2IError
    mask = torch.ones(3, 9, dtype=torch.long)
    mask[1, :4] = 0
    mask[2, :8] = 0
    pos, start = mllm.Qwen2_5_VLTextPrefill.text_positions(mask)
    assert torch.equal(pos.long(), mo.text_positions(mask)[0]) and pos.dtype == torch.int32
    assert start.tolist() == [0, 4, 8]
    assert pos[1].tolist() == [1, 1, 1, 1, 0, 1, 2, 3, 4]
    bad = torch.ones(1, 6, dtype=torch.long)
    bad[0, 4:] = 0  # right padding
    with pytest.raises(X2IError):
        mllm.Qwen2_5_VLTextPrefill.text_positions(bad)


def test_other_mllm_families_share_the_decoder():
    """InternVL2.5-4B (Qwen2.5-3B-Instruct inside) and MiniCPM-o-2.6 (Qwen2.5-7B inside) -- the MLLMs of infer/inference_internvl.py
    and infer/inference_minicpm.py -- run the same text decoder: their configurations produce the projectors' C x H (37 x 2048,
    29 x 3584: utils/proj.py create_proj_internvl4b / create_proj_minicpm), a plain ``Qwen2Model`` with the same weights gives the
    same states as the Qwen2.5-VL text model, their checkpoint prefixes load, and the two position rules (a forward without
    position_ids = arange; HF generate = cumsum(mask) - 1) differ only by rounding at the real tokens (RoPE is relative)."""
    from oracle import mllm_oracle as mo
    from x2i_b200 import mllm
    assert (mllm.INTERNVL2_5_4B_LLM["num_hidden_layers"] + 1, mllm.INTERNVL2_5_4B_LLM["hidden_size"]) == (37, 2048)
    assert (mllm.MINICPM_O_2_6_LLM["num_hidden_layers"] + 1, mllm.MINICPM_O_2_6_LLM["hidden_size"]) == (29, 3584)
    assert mllm.INTERNVL2_5_4B_LLM["position_mode"] == "arange" and mllm.MINICPM_O_2_6_LLM["position_mode"] == "cumsum"
    plain = mo.build_qwen2(TINY, seed=1)
    vl = mo.build(TINY, seed=2)
    vl.load_state_dict(plain.state_dict(), strict=False)
    ids = torch.randint(0, 1000, (2, 40), generator=torch.Generator().manual_seed(3))
    mask = torch.ones(2, 40, dtype=torch.long)
    mask[1, :13] = 0
    a = mo.prefill_hidden_states_plain(plain, ids, mask, "arange")
    c = mo.prefill_hidden_states_plain(plain, ids, mask, "cumsum")
    v = mask.bool()[:, None, :, None].expand_as(a)
    assert torch.equal(c, mo.prefill_hidden_states(vl, ids, mask))                # same weights, same rule: the same model
    assert float((a[v] - c[v]).norm() / c[v].norm()) < 1e-5                        # arange vs cumsum: rounding only
    pos, start = mllm.Qwen2_5_VLTextPrefill.text_positions(mask, "arange")
    assert pos[1].tolist() == list(range(40)) and start.tolist() == [0, 13]
    m = mllm.Qwen2_5_VLTextPrefill(**dict(TINY, position_mode="arange"))
    for prefix in ("language_model.model.", "llm.model."):                         # InternVLChatModel / MiniCPMO checkpoints
        sd = {prefix + k: v_ for k, v_ in plain.state_dict().items()}
        sd.update({"vision_model.embeddings.class_embedding": torch.zeros(1), "mlp1.0.weight": torch.zeros(1), "vpm.x": torch.zeros(1),
                   "resampler.query": torch.zeros(1), "language_model.lm_head.weight": torch.zeros(1), "llm.lm_head.weight": torch.zeros(1)})
        m.load_hf_state_dict(sd)
        assert torch.equal(m.layers[1].self_attn.k_proj.bias, plain.layers[1].self_attn.k_proj.bias)


def test_swiglu_weight_packing_layout():
    from x2i_b200 import ops
    g = torch.arange(256 * 8, dtype=torch.float32).view(256, 8)
    u = -g
    w = ops.pack_swiglu_weight(g, u)
    assert w.shape == (512, 8)
    assert torch.equal(w[:128], g[:128]) and torch.equal(w[128:256], u[:128]) and torch.equal(w[256:384], g[128:]) and torch.equal(w[384:], u[128:])


def test_oracle_captures_like_the_reference_call():
    """hidden_states of the prefill step: embeddings, layer outputs, and the final-normed last layer; stack(dim=1) == cat().unsqueeze(0) for B=1."""
    from oracle import mllm_oracle as mo
    o = mo.build(TINY)
    ids = torch.randint(0, 1000, (1, 12), generator=torch.Generator().manual_seed(1))
    mask = torch.ones(1, 12, dtype=torch.long)
    te = mo.prefill_hidden_states(o, ids, mask)
    assert te.shape == (1, 4, 12, 256)
    with torch.no_grad():
        out = o(input_ids=ids, attention_mask=mask, output_hidden_states=True)
    assert torch.equal(torch.cat(out.hidden_states).unsqueeze(0), te)          # infer/inference_qwenvl.py:123
    assert torch.equal(te[:, 0], o.embed_tokens(ids)) and torch.equal(te[:, -1], out.last_hidden_state)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import ops
    return ops


def rn(*s, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*s, device="cuda", generator=g) * scale).bfloat16()


@gpu
@pytest.mark.parametrize("B,H,Hkv,L,starts", [(2, 4, 2, 300, (0, 37)), (1, 16, 2, 512, (0,)), (3, 2, 1, 130, (0, 129, 64)), (2, 8, 1, 700, (300, 0))])
def test_causal_gqa_attention_with_left_padding(ops, B, H, Hkv, L, starts):
    q, k, v = rn(B, H, L, 128, seed=1), rn(B, Hkv, L, 128, seed=2), rn(B, Hkv, L, 128, seed=3)
    st = torch.tensor(starts, device="cuda", dtype=torch.int32)
    out = ops.causal_attention(q, k, v, kv_start=st).view(B, L, H, 128).transpose(1, 2)
    i = torch.arange(L, device="cuda")
    vis = (i[None, :, None] >= i[None, None, :]) & (i[None, None, :] >= st[:, None, None].long())  # [B, Lq, Lk]
    kk, vv = k.float().repeat_interleave(H // Hkv, 1), v.float().repeat_interleave(H // Hkv, 1)
    s = (q.float() @ kk.transpose(-1, -2)) / 128 ** 0.5
    s = s.masked_fill(~vis[:, None], float("-inf"))
    p = torch.softmax(s, -1).nan_to_num(0.0)       # rows with no visible key -> 0 (sdpa / flash semantics)
    ref = p @ vv
    assert rel(out, ref) < 6e-3
    for b, s0 in enumerate(starts):
        if s0 > 0:
            assert float(out[b, :, :s0].float().abs().max()) == 0.0  # padded query rows output exactly zero
    assert torch.equal(ops.causal_attention(q, k, v, kv_start=st).view(B, L, H, 128).transpose(1, 2), out)  # deterministic
    import os
    os.environ["X2I_ATTN_PERSIST"] = "1" if os.environ.get("X2I_ATTN_PERSIST", "") != "1" else "0"  # the other kernel form: bit-identical
    try:
        assert torch.equal(ops.causal_attention(q, k, v, kv_start=st).view(B, L, H, 128).transpose(1, 2), out)
    finally:
        os.environ.pop("X2I_ATTN_PERSIST", None)


@gpu
def test_prefill_rowwise_kernels_and_swiglu_gemm(ops):
    import torch.nn.functional as F
    # embedding gather into a strided layer slot: bit-exact
    table = rn(500, 256, seed=4)
    ids = torch.randint(0, 500, (2, 33), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    buf = torch.zeros(2, 3, 33, 256, device="cuda", dtype=torch.bfloat16)
    ops.gather_rows(ids, table, buf[:, 1])
    assert torch.equal(buf[:, 1], table[ids]) and float(buf[:, 0].abs().max()) == 0 and float(buf[:, 2].abs().max()) == 0
    # RMSNorm (Qwen2RMSNorm rounding order) on a batch-strided view
    w = (1 + 0.1 * rn(256, seed=6).float()).bfloat16()
    y = ops.rmsnorm(buf[:, 1], w, 1e-6)
    x = buf[:, 1].float()
    ref = w.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
    assert rel(y, ref) < 4e-3
    for D in (2048, 3584):
        xx, ww = rn(3, 17, D, seed=7), (1 + 0.1 * rn(D, seed=8).float()).bfloat16()
        r2 = ww.float() * (xx.float() * torch.rsqrt(xx.float().pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float()
        assert rel(ops.rmsnorm(xx, ww, 1e-6), r2) < 4e-3
    # rotate-half RoPE + head-major split vs transformers' apply_rotary_pos_emb arithmetic in fp32
    B, S, Hq, Hkv = 2, 50, 4, 2
    qkv = rn(B, S, (Hq + 2 * Hkv) * 128, seed=9)
    pos = torch.randint(0, 600, (B, S), device="cuda", generator=torch.Generator(device="cuda").manual_seed(10)).int()
    inv_freq = (1.0 / (1e6 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))).cuda()
    q, k, v = ops.rope_half_split(qkv, pos, inv_freq, Hq, Hkv)
    fr = pos.float()[..., None] * inv_freq  # [B, S, 64]
    cos, sin = torch.cat([fr, fr], -1).cos()[:, None], torch.cat([fr, fr], -1).sin()[:, None]
    def rot(t):
        return torch.cat([-t[..., 64:], t[..., :64]], -1)
    qr = qkv[..., :Hq * 128].float().view(B, S, Hq, 128).transpose(1, 2)
    kr = qkv[..., Hq * 128:(Hq + Hkv) * 128].float().view(B, S, Hkv, 128).transpose(1, 2)
    vr = qkv[..., (Hq + Hkv) * 128:].view(B, S, Hkv, 128).transpose(1, 2)
    assert rel(q, qr * cos + rot(qr) * sin) < 4e-3 and rel(k, kr * cos + rot(kr) * sin) < 4e-3 and torch.equal(v, vr)
    # SwiGLU GEMM: CTA-pair kernel (M > 128) and single-CTA kernel (M <= 128)
    for M in (300, 100):
        xg, wg, wu = rn(M, 256, seed=11), rn(384, 256, seed=12, scale=0.06), rn(384, 256, seed=13, scale=0.06)
        got = ops.linear_swiglu(xg, ops.pack_swiglu_weight(wg, wu))
        want = F.silu(xg.float() @ wg.float().T) * (xg.float() @ wu.float().T)
        assert got.shape == (M, 384) and rel(got, want) < 4e-3
    # plain residual GEMM (gate = NULL) into a strided destination
    xr, wr, res = rn(200, 384, seed=14), rn(256, 384, seed=15, scale=0.05), rn(200, 256, seed=16)
    dst = torch.zeros(200, 512, device="cuda", dtype=torch.bfloat16)
    ops.linear_residual(xr, wr, res, dst[:, :256])
    assert rel(dst[:, :256], res.float() + xr.float() @ wr.float().T) < 4e-3 and float(dst[:, 256:].abs().max()) == 0


def _pair(cfg, seed, device="cuda"):
    from oracle import mllm_oracle as mo
    from x2i_b200 import mllm
    o = mo.build(cfg, seed=seed)
    with torch.no_grad():
        for p in o.parameters():
            p.copy_((p if p.ndim < 2 else p * 1.0).bfloat16().float())
    m = mllm.Qwen2_5_VLTextPrefill(**cfg)
    m.load_hf_state_dict(o.state_dict())
    return m.to(device, torch.bfloat16).eval(), o


@gpu
@pytest.mark.parametrize("B,S,pads", [(2, 40, (0, 13)), (2, 300, (0, 170)), (1, 512, (0,))])
def test_tiny_prefill_matches_transformers(ops, B, S, pads):
    from oracle import mllm_oracle as mo
    m, o = _pair(TINY, seed=3)
    ids = torch.randint(0, 1000, (B, S), generator=torch.Generator().manual_seed(4))
    mask = torch.ones(B, S, dtype=torch.long)
    for b, p in enumerate(pads):
        mask[b, :p] = 0
    ref = mo.prefill_hidden_states(o, ids, mask)                                             # fp32, CPU
    got = m.prefill_hidden_states(ids.cuda(), mask.cuda())
    eag = mo.prefill_hidden_states(o.to("cuda", torch.bfloat16), ids.cuda(), mask.cuda())    # the reference's own dtype / backend
    assert got.shape == ref.shape == (B, 4, S, 256) and got.dtype == torch.bfloat16
    assert torch.equal(got[:, 0].cpu().float(), ref[:, 0])                                   # embeddings: bit-exact index work
    valid = mask.bool()[:, None, :, None].expand_as(ref)
    check(f"tiny MLLM prefill B={B} S={S} pads={pads}: all layers, valid tokens", rel(got.cpu().float()[valid], ref[valid]),
          rel(eag.cpu().float()[valid], ref[valid]))
    if any(pads):  # padded positions: zero attention output (sdpa semantics) -- same states as the library's sdpa path
        check(f"tiny MLLM prefill B={B} S={S}: padded positions", rel(got.cpu().float()[~valid], ref[~valid]), rel(eag.cpu().float()[~valid], ref[~valid]))
    out = m.generate(input_ids=ids.cuda(), attention_mask=mask.cuda(), max_new_tokens=1, output_hidden_states=True, return_dict_in_generate=True)
    assert torch.equal(torch.stack(out["hidden_states"][0], dim=1), got)                    # train_qwenvl.py:775 on the drop-in's output
    assert out.text_embeddings.data_ptr() == out["hidden_states"][0][0].data_ptr()          # the tuple entries are views: no copy
    # the default path replays one captured CUDA graph per (B, S); launch by launch it computes the same bits, and a second prompt of the
    # same shape through the same graph is not contaminated by the first
    m.use_cuda_graph = False
    try:
        assert torch.equal(m.prefill_hidden_states(ids.cuda(), mask.cuda()), got)
        ids2 = torch.randint(0, 1000, (B, S), generator=torch.Generator().manual_seed(5))
        eager2 = m.prefill_hidden_states(ids2.cuda(), mask.cuda())
    finally:
        m.use_cuda_graph = True
    assert torch.equal(m.prefill_hidden_states(ids2.cuda(), mask.cuda()), eager2)
    user_out = torch.empty_like(got)
    assert m.prefill_hidden_states(ids.cuda(), mask.cuda(), out=user_out) is user_out and torch.equal(user_out, got)


@gpu
def test_internvl_style_prefill_matches_plain_qwen2_forward(ops):
    """The InternVL path (model_internvl/internvl/modeling_internvl_chat.py:357-363): a plain Qwen2 language model called without
    position_ids on a left-padded prompt -> positions = arange(S)."""
    from oracle import mllm_oracle as mo
    from x2i_b200 import mllm
    o = mo.build_qwen2(TINY, seed=21)
    with torch.no_grad():
        for p in o.parameters():
            p.copy_(p.bfloat16().float())
    m = mllm.Qwen2_5_VLTextPrefill(**dict(TINY, position_mode="arange"))
    m.load_hf_state_dict({"language_model.model." + k: v for k, v in o.state_dict().items()})
    m = m.to("cuda", torch.bfloat16).eval()
    B, S = 2, 300
    ids = torch.randint(0, 1000, (B, S), generator=torch.Generator().manual_seed(22))
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, :170] = 0
    ref = mo.prefill_hidden_states_plain(o, ids, mask, "arange")
    got = m.prefill_hidden_states(ids.cuda(), mask.cuda())
    eag = mo.prefill_hidden_states_plain(o.to("cuda", torch.bfloat16), ids.cuda(), mask.cuda(), "arange")
    valid = mask.bool()[:, None, :, None].expand_as(ref)
    check("InternVL-style prefill (plain Qwen2, arange positions): all layers, valid tokens", rel(got.cpu().float()[valid], ref[valid]),
          rel(eag.cpu().float()[valid], ref[valid]))


@gpu
def test_qwen2_5_vl_3b_prefill_full_size_feeds_the_projector(ops):
    """BASELINE config 2's conditioning path at full size: Qwen2.5-VL-3B text decoder (36 layers, 2048 wide, 16/2 heads, random weights),
    prompts padded to 512 on the left, all 37 hidden states -> [B, 37, 512, 2048] -> Proj7Exp (qwen3b, use_cnn).  fp32 oracle on the GPU."""
    from oracle import mllm_oracle as mo, proj_oracle
    from x2i_b200 import mllm, proj as xproj
    from x2i_b200.flux import init_synthetic_
    cfg = mllm.QWEN2_5_VL_3B
    o = mo.build(dict(cfg, vocab_size=8192), seed=5, empty_on="cuda")  # smaller vocabulary: the fp32 oracle stays at 11 GB; ids are drawn below it
    init_synthetic_(o, seed=6, std=0.02)
    with torch.no_grad():
        for p in o.parameters():
            p.copy_(p.bfloat16().float())
    m = mllm.Qwen2_5_VLTextPrefill(**dict(cfg, vocab_size=8192))
    m.load_hf_state_dict(o.state_dict())
    m = m.to("cuda", torch.bfloat16).eval()
    B, S = 2, 512
    ids = torch.randint(0, 8192, (B, S), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    mask = torch.ones(B, S, dtype=torch.long, device="cuda")
    mask[1, :389] = 0  # a 123-token prompt, left-padded to 512 like the reference's processor call
    ref = mo.prefill_hidden_states(o, ids, mask)
    got = m.prefill_hidden_states(ids, mask)
    eag = mo.prefill_hidden_states(o.to(torch.bfloat16), ids, mask)
    assert got.shape == (B, 37, S, 2048)
    valid = mask.bool()[:, None, :, None].expand_as(ref)
    per_layer = [(rel(got[:, c][valid[:, c]], ref[:, c][valid[:, c]]), rel(eag[:, c][valid[:, c]], ref[:, c][valid[:, c]])) for c in range(37)]
    record("Qwen2.5-VL-3B prefill: worst layer (valid tokens)", max(e for e, _ in per_layer), max(y for _, y in per_layer))
    for c, (e, y) in enumerate(per_layer):
        assert e <= 1e-2 or e <= y, f"layer {c}: rel err {e:.5f} (eager bf16 {y:.5f})"
    check("Qwen2.5-VL-3B prefill [2,37,512,2048]: all layers, valid tokens", rel(got[valid], ref[valid]), rel(eag[valid], ref[valid]))
    check("Qwen2.5-VL-3B prefill: padded positions (sdpa semantics)", rel(got[~valid], ref[~valid]), rel(eag[~valid], ref[~valid]))
    # projector on top (the reference's next line: proj(text_embeddings), inference_qwenvl.py:179)
    po = proj_oracle.Proj7Exp(in_channels=37, input_dim=2048, use_scale=False, use_cnn=True).cuda()
    init_synthetic_(po, seed=8, std=0.02)
    with torch.no_grad():
        for p in po.parameters():
            p.copy_(p.bfloat16().float())
    pm = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True)
    pm.load_state_dict(po.state_dict())
    pm = pm.to("cuda", torch.bfloat16)
    with torch.no_grad():
        rp, rs = po(ref)
        gp, gs = pm(got)
        ep, es = po.to(torch.bfloat16)(eag)
    check("prefill -> projector: prompt_embeds [2,512,4096]", rel(gs, rs), rel(es, rs))
    check("prefill -> projector: pooled [2,768]", rel(gp, rp), rel(ep, rp))


@gpu
def test_qwen2_5_vl_7b_width_prefill_feeds_the_7b_projector(ops):
    """The 7B variant of the conditioning path (infer/inference_qwenvl.py:80-81, train_qwenvl.py:401: `create_proj3_qwen7b(in_channels=29)`)
    at its real WIDTH and reduced depth: hidden 3584, 28 query / 4 key-value heads (7 : 1 grouped-query attention), intermediate 18944,
    28 layers in the checkpoint -- 6 here, so the projector is built for C = 7.  Covers the K = 3584 GEMM shapes, the 28 / 4 head split of the
    RoPE kernel and the causal attention kernel, and the 3584-wide layer-mixing convolution."""
    from oracle import mllm_oracle as mo, proj_oracle
    from x2i_b200 import mllm, proj as xproj
    from x2i_b200.flux import init_synthetic_
    cfg = dict(mllm.QWEN2_5_VL_7B, vocab_size=4096, num_hidden_layers=6)
    o = mo.build(cfg, seed=15, empty_on="cuda")
    init_synthetic_(o, seed=16, std=0.02)
    with torch.no_grad():
        for p in o.parameters():
            p.copy_(p.bfloat16().float())
    m = mllm.Qwen2_5_VLTextPrefill(**cfg)
    m.load_hf_state_dict(o.state_dict())
    m = m.to("cuda", torch.bfloat16).eval()
    B, S = 2, 512
    ids = torch.randint(0, 4096, (B, S), device="cuda", generator=torch.Generator(device="cuda").manual_seed(17))
    mask = torch.ones(B, S, dtype=torch.long, device="cuda")
    mask[0, :200] = 0
    ref = mo.prefill_hidden_states(o, ids, mask)
    got = m.prefill_hidden_states(ids, mask)
    eag = mo.prefill_hidden_states(o.to(torch.bfloat16), ids, mask)
    assert got.shape == (B, 7, S, 3584)
    valid = mask.bool()[:, None, :, None].expand_as(ref)
    check("Qwen2.5-VL-7B width (6 layers) prefill [2,7,512,3584]: all layers, valid tokens", rel(got[valid], ref[valid]), rel(eag[valid], ref[valid]))
    po = proj_oracle.Proj7Exp(in_channels=7, input_dim=3584, use_scale=False, use_cnn=True).cuda()
    init_synthetic_(po, seed=18, std=0.02)
    with torch.no_grad():
        for p in po.parameters():
            p.copy_(p.bfloat16().float())
    pm = xproj.create_proj3_qwen7b(7, use_t5=False, use_scale=False, use_cnn=True)
    pm.load_state_dict(po.state_dict())
    pm = pm.to("cuda", torch.bfloat16)
    with torch.no_grad():
        rp, rs = po(ref)
        gp, gs = pm(got)
        ep, es = po.to(torch.bfloat16)(eag)
    check("7B-width prefill -> projector (qwen7b): prompt_embeds [2,512,4096]", rel(gs, rs), rel(es, rs))
    check("7B-width prefill -> projector (qwen7b): pooled [2,768]", rel(gp, rp), rel(ep, rp))
