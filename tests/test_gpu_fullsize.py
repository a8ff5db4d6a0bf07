"""-m gpu: the hot-path kernels at BASELINE.json's FULL sizes, checked through size-independent properties
(the fp32 oracle needs minutes per block at these sizes, so the small-shape parity tests carry the element-wise
comparison and these carry the scale).

Shapes: attention (B, 24 heads, 4096 + 512 tokens, d 128) = configs 2/3 at 1024 px; KD rows = 4608 x 3072 per layer
(train/train_qwenvl.py:601-620); projector input [1, 37, 512, 2048] (infer/inference_qwenvl.py:80).
Tolerances: bf16 outputs -> 1e-2 relative, exact where the property is exact.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

H, L_TXT, L_IMG, DH = 24, 512, 4096, 128
L = L_TXT + L_IMG


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import ops
    return ops


def _qkv(B, seed, scale=1.0, outliers=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q, k, v = (torch.randn(B, H, L, DH, device="cuda", generator=g) * scale for _ in range(3))
    if outliers:  # 4 random channels x40 (FLUX activations carry such channels)
        ch = torch.randperm(DH, device="cuda", generator=g)[:4]
        q[..., ch] *= 40.0
        k[..., ch] *= 40.0
    return q.bfloat16(), k.bfloat16(), v.bfloat16()


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _attn(ops, q, k, v):
    B = q.shape[0]
    o_txt = torch.empty(B, L_TXT, H * DH, device="cuda", dtype=torch.bfloat16)
    o_img = torch.empty(B, L_IMG, H * DH, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, split=L_TXT, out0=o_txt, out1=o_img)
    return torch.cat([o_txt, o_img], 1).view(B, L, H, DH).transpose(1, 2)  # [B, H, L, DH]


@pytest.mark.parametrize("outliers", [False, True])
def test_attention_full_size_rows_are_convex_combinations(ops, outliers):
    """softmax rows sum to one: with V == const the output is that constant; with V = one-hot(d) per key parity the
    output channels are probabilities in [0, 1] that sum to one."""
    q, k, v = _qkv(1, 11, outliers=outliers)
    ones = torch.full_like(v, 0.75)
    o = _attn(ops, q, k, ones)
    assert torch.isfinite(o.float()).all()
    assert float((o.float() - 0.75).abs().max()) <= 0.75 * 2 ** -7  # one bf16 ulp of the P rounding + output rounding
    e = torch.zeros_like(v)
    e[:, :, 0::2, 0] = 1.0
    e[:, :, 1::2, 1] = 1.0
    o = _attn(ops, q, k, e).float()
    s = o[..., 0] + o[..., 1]
    assert float((s - 1.0).abs().max()) < 1e-2
    assert float(o.min()) >= 0.0 and float(o[..., 2:].abs().max()) == 0.0


def test_attention_full_size_linear_in_v_and_key_permutation_invariant(ops):
    q, k, v = _qkv(2, 12)
    g = torch.Generator(device="cuda").manual_seed(5)
    v2 = torch.randn(v.shape, device="cuda", generator=g).bfloat16()
    o1, o2 = _attn(ops, q, k, v).float(), _attn(ops, q, k, v2).float()
    o12 = _attn(ops, q, k, (v.float() * 0.5 + v2.float() * 0.25).bfloat16()).float()
    assert _rel(o12, 0.5 * o1 + 0.25 * o2) < 1e-2
    perm = torch.randperm(L, device="cuda", generator=g)
    op = _attn(ops, q, k[:, :, perm].contiguous(), v[:, :, perm].contiguous()).float()
    assert _rel(op, o1) < 6e-3  # different tile order -> different rounding only


def test_attention_full_size_matches_fp32_on_sampled_heads(ops):
    """Element-wise against torch fp32 on 3 of the 24 heads (the full fp32 reference would need 2 GB of scores per head)."""
    q, k, v = _qkv(1, 13, outliers=True)
    o = _attn(ops, q, k, v)
    for h in (0, 11, 23):
        ref = torch.nn.functional.scaled_dot_product_attention(q[:, h:h + 1].float(), k[:, h:h + 1].float(), v[:, h:h + 1].float())
        assert _rel(o[:, h:h + 1], ref) < 6e-3


def test_attention_kernel_forms_are_bit_identical(ops):
    """The one-item kernel, the persistent kernel (attn_persist_sm100.cuh: one CTA per SM loops over work items, X2I_ATTN_PERSIST) and the
    CTA-pair kernel (attn2_sm100.cuh: cta_group::2, two CTAs share every K / V tile, X2I_ATTN_PAIR) run the same soft-max code on the same
    tiles: bit-identical outputs at the full size, on a ragged length (odd number of query blocks, partial last key tile) and with more
    work items than SMs can hold in one round."""
    import os
    # X2I_ATTN_LAG=0: the classic soft-max step in the persistent kernel (its default, the lagged step, is not bit-identical to the other forms:
    # test_attention_lagged_form_matches_fp32_and_default)
    modes = [dict(X2I_ATTN_PAIR="0", X2I_ATTN_PERSIST="0"), dict(X2I_ATTN_PAIR="0", X2I_ATTN_PERSIST="1", X2I_ATTN_LAG="0"),
             dict(X2I_ATTN_PAIR="1", X2I_ATTN_PERSIST="0")]
    for Bq, Hq, Lq in ((1, 4, L), (1, 4, 600), (2, 24, 1100)):
        g = torch.Generator(device="cuda").manual_seed(17)
        q, k, v = (torch.randn(Bq, Hq, Lq, DH, device="cuda", generator=g).bfloat16() for _ in range(3))
        outs = []
        for mode in modes:
            os.environ.update(mode)
            try:
                o = torch.empty(Bq, Lq, Hq * DH, device="cuda", dtype=torch.bfloat16)
                ops.attention(q, k, v, split=0, out1=o)
                outs.append(o)
            finally:
                for kk in mode:
                    os.environ.pop(kk, None)
        assert torch.equal(outs[0], outs[1]), "persistent kernel differs"
        assert torch.equal(outs[0], outs[2]), "CTA-pair kernel differs"
        ref = torch.nn.functional.scaled_dot_product_attention(q[:, :1].float(), k[:, :1].float(), v[:, :1].float())
        assert _rel(outs[1].view(Bq, Lq, Hq, DH).transpose(1, 2)[:, :1], ref) < 6e-3
        # column-split soft-max (attn_cs_sm100.cuh: both warpgroups on the same query tile, row max / row sum combined from two halves):
        # same exponentials and the same P, the row sum is added in a different order -> equal up to the last bf16 bit of O
        os.environ["X2I_ATTN_CS"] = "1"
        try:
            ocs = torch.full((Bq, Lq, Hq * DH), float("nan"), device="cuda", dtype=torch.bfloat16)
            ops.attention(q, k, v, split=0, out1=ocs)
        finally:
            os.environ.pop("X2I_ATTN_CS", None)
        assert _rel(ocs.view(Bq, Lq, Hq, DH).transpose(1, 2)[:, :1], ref) < 6e-3
        assert _rel(ocs, outs[1]) < 2e-3 and float((ocs.float() - outs[1].float()).abs().max()) < 0.05
        os.environ.update(X2I_ATTN_CS="0", X2I_ATTN_LAG="0")
        try:
            o0 = torch.empty_like(ocs)
            ops.attention(q, k, v, split=0, out1=o0)
        finally:
            os.environ.pop("X2I_ATTN_CS", None)
            os.environ.pop("X2I_ATTN_LAG", None)
        assert torch.equal(o0, outs[1])


def test_attention_lagged_form_matches_fp32_and_default(ops):
    """Lagged soft-max steps (softmax_step_lagged, X2I_ATTN_LAG=1): the exponentials of a key step run against the reference the row
    already has, the next reference comes from the step's row sum (one step late); a step whose sum exceeds 2^96 -- or whose
    polynomial-lane arguments exceed 126 -- makes the CTA re-run its items with the classic step.  Checked against torch fp32 and the
    default form on (a) random scores, (b) scores that GROW along the keys by ~2^42 per tile (a lagged rescale at every step), (c) the
    same with ~2^167 per tile (MUFU lanes return inf -> redo pass), (d) a few late keys that dominate single rows by ~2^25, (e) ONE late
    key ~2^200 above everything else, once on a polynomial lane of the soft-max (its exponent insertion would wrap around silently) and
    once on a MUFU lane, at the full size, a ragged length and more work items than SMs."""
    import os

    def run(q, k, v, lag):
        os.environ["X2I_ATTN_LAG"] = "1" if lag else "0"
        try:
            Bq, Hq, Lq, _ = q.shape
            o = torch.full((Bq, Lq, Hq * DH), float("nan"), device="cuda", dtype=torch.bfloat16)
            ops.attention(q, k, v, split=0, out1=o)
            return o
        finally:
            os.environ.pop("X2I_ATTN_LAG", None)

    for Bq, Hq, Lq in ((1, 4, L), (1, 4, 600), (2, 24, 1100)):
        g = torch.Generator(device="cuda").manual_seed(23)
        q, k, v = (torch.randn(Bq, Hq, Lq, DH, device="cuda", generator=g) for _ in range(3))
        cases = {"random": (q, k)}
        # q.k_j = 4 * 128 * slope * j: slope 0.005 -> +328 per 128 keys in raw score units = +42 in the exp2 domain; 0.02 -> +167
        for name, slope in (("ramp", 0.005), ("steep ramp", 0.02)):
            ramp = (torch.arange(Lq, device="cuda", dtype=torch.float32) * slope).view(1, 1, Lq, 1)
            cases[name] = (torch.full_like(q, 4.0) + 0.1 * q, ramp.expand(Bq, Hq, Lq, DH) + 0.05 * k)
        # 16 late keys aligned with 16 query rows: a spike of ~ +25 in the exp2 domain, elsewhere random
        k2 = k.clone()
        rows = torch.randperm(Lq, device="cuda", generator=g)[:16]
        keys = torch.randint(Lq // 2, Lq, (16,), device="cuda", generator=g)
        k2[:, :, keys] = q[:, :, rows] * 1.5
        cases["spikes"] = (q, k2)
        # one key far above everything: q = 3 + noise, that key = 4 -> q.k ~ 1536 raw = +196 in the exp2 domain.  Lane of key index t within
        # its 128-key tile: pair (t % 32) // 2; pairs 0, 1, 8, 9 of every quarter run on the polynomial (POLY8 = 2), the rest on MUFU.EX2
        for name, t_in_tile in (("huge key on a polynomial lane", 32 + 2), ("huge key on a MUFU lane", 32 + 10)):
            k3 = k.clone()
            k3[:, :, 256 + t_in_tile] = 4.0
            cases[name] = (torch.full_like(q, 3.0) + 0.1 * q, k3)
        for name, (qq, kk) in cases.items():
            qb, kb, vb = qq.bfloat16(), kk.bfloat16(), v.bfloat16()
            o_lag, o_def = run(qb, kb, vb, True), run(qb, kb, vb, False)
            assert torch.isfinite(o_lag.float()).all(), (name, Lq)
            ref = torch.nn.functional.scaled_dot_product_attention(qb[:, :1].float(), kb[:, :1].float(), vb[:, :1].float())
            e_lag = _rel(o_lag.view(Bq, Lq, Hq, DH).transpose(1, 2)[:, :1], ref)
            e_def = _rel(o_def.view(Bq, Lq, Hq, DH).transpose(1, 2)[:, :1], ref)
            assert e_lag < 6e-3, (name, Lq, e_lag, e_def)
            assert _rel(o_lag, o_def) < 6e-3, (name, Lq)


def test_attention_full_size_shift_invariance(ops):
    """softmax(s + c) == softmax(s): adding a constant vector to every key along a direction orthogonal to nothing
    changes all scores of a row by the same amount q.c -> same output (up to bf16 rounding of k)."""
    q, k, v = _qkv(1, 14)
    c = torch.zeros(DH, device="cuda")
    c[3] = 8.0  # exactly representable shift of one channel
    o1 = _attn(ops, q, k, v).float()
    o2 = _attn(ops, q, (k.float() + c).bfloat16(), v).float()
    assert _rel(o2, o1) < 1e-2


def test_kd_loss_full_size_properties(ops):
    """One FLUX layer pair [1, 4608, 3072] (x 4 layers): loss(x, x) == 0 with zero gradient, loss is invariant under a per-row
    affine map of either argument (normalize(), train_qwenvl.py:58-61), positive otherwise, and the gradient is
    orthogonal to the per-row affine directions (1 and x)."""
    from x2i_b200 import kd
    g = torch.Generator(device="cuda").manual_seed(3)
    t = torch.randn(1, 4, L, 3072, device="cuda", generator=g).bfloat16()  # [B, n_layers, rows, D]
    s = torch.randn(1, 4, L, 3072, device="cuda", generator=g).bfloat16().requires_grad_(True)
    same = kd.kd_loss_stacked(t, t.clone().requires_grad_(True))[0]
    assert abs(float(same.detach())) < 1e-3
    loss, terms, valid = kd.kd_loss_stacked(t, s)
    assert valid.tolist() == [1, 1, 1, 1]
    loss.backward()
    assert float(loss) > 0 and torch.isfinite(s.grad.float()).all()
    a = torch.rand(1, 4, L, 1, device="cuda", generator=g) * 3 + 0.5
    b = torch.randn(1, 4, L, 1, device="cuda", generator=g)
    loss_aff = kd.kd_loss_stacked((t.float() * a + b).bfloat16(), s.detach())[0]
    assert abs(float(loss_aff) - float(loss)) / float(loss) < 1e-2
    gr, x = s.grad.float(), s.detach().float()
    scale = gr.abs().sum(-1) + 1e-20
    assert float((gr.sum(-1).abs() / scale).max()) < 2e-2          # d loss / d shift == 0
    assert float(((gr * x).sum(-1).abs() / (scale * x.abs().amax(-1))).max()) < 2e-2  # d loss / d scale == 0


def test_projector_full_size_batch_and_sequence_independence(ops):
    """Proj7Exp at config-2 size [B, 37, 512, 2048]: samples are independent (batch of 2 == two batches of 1, bit-exact) and
    the per-token branch is independent of the other tokens (a changed token changes only its own row of the sequence
    output; the pooled output is the mean over tokens)."""
    from x2i_b200 import proj as xproj
    torch.manual_seed(0)
    m = xproj.create_proj3_qwen3b(37, use_t5=False, use_scale=False, use_cnn=True).to("cuda", torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(2, 37, 512, 2048, device="cuda", generator=g).bfloat16()
    with torch.no_grad():
        p2, s2 = m(x)
        p0, s0 = m(x[:1].contiguous())
        p1, s1 = m(x[1:].contiguous())
    assert p2.shape == (2, 768) and s2.shape == (2, 512, 4096)
    assert torch.equal(s2[0], s0[0]) and torch.equal(s2[1], s1[0])
    assert torch.isfinite(s2.float()).all() and torch.isfinite(p2.float()).all()
    assert _rel(p2[0], p0[0]) < 1e-2


def test_denoise_step_full_size_batch_consistency_and_determinism():
    """One full-width 1024 px denoise step on a reduced-depth model (2 double + 2 single blocks, all other dimensions as in
    FLUX): sample b of a batch of 2 equals the same sample run alone, and two runs are bit-identical."""
    from x2i_b200.flux import FluxTransformer2DModel
    from x2i_b200.pipeline import FluxPipeline
    cfg = dict(patch_size=1, in_channels=64, num_layers=2, num_single_layers=2, attention_head_dim=128, num_attention_heads=24,
               joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    dev = torch.device("cuda")
    model = FluxTransformer2DModel.synthetic(cfg, device=dev, seed=0)
    model.use_cuda_graph = False
    g = torch.Generator(device=dev).manual_seed(21)
    B = 2
    lat = torch.randn(B, L_IMG, 64, device=dev, generator=g).bfloat16()
    prompt = torch.randn(B, L_TXT, 4096, device=dev, generator=g).bfloat16()
    pooled = torch.randn(B, 768, device=dev, generator=g).bfloat16()
    img_ids = FluxPipeline._prepare_latent_image_ids(B, 128, 128, dev, torch.bfloat16)
    txt_ids = torch.zeros(L_TXT, 3, device=dev, dtype=torch.bfloat16)

    def run(sl):
        n = lat[sl].shape[0]
        t = torch.full((n,), 0.75, device=dev, dtype=torch.bfloat16)
        gd = torch.full((n,), 3.5, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            return model(hidden_states=lat[sl].contiguous(), timestep=t, guidance=gd, pooled_projections=pooled[sl].contiguous(),
                         encoder_hidden_states=prompt[sl].contiguous(), txt_ids=txt_ids, img_ids=img_ids, return_dict=False)[0].clone()

    both, again = run(slice(0, 2)), run(slice(0, 2))
    assert torch.equal(both, again)
    assert torch.isfinite(both.float()).all() and float(both.float().abs().mean()) > 0
    for b in range(B):
        alone = run(slice(b, b + 1))
        assert _rel(both[b:b + 1], alone) < 1e-2  # tile scheduling differs with B; arithmetic per sample is the same
