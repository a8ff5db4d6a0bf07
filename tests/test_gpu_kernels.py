"""-m gpu: every C-ABI kernel against a plain torch fp32 computation of the same op on the same bf16 inputs.

Tolerances (written here, per BASELINE.md: <= 1e-2 relative in bf16; bit-exact for index/gather work):
  * one bf16 rounding of the output gives a relative Frobenius error of ~1.7e-3 -> bound 4e-3 for single ops,
    6e-3 for the attention kernel (P is rounded to bf16 before the PV contraction);
  * RoPE table / Euler update / latent ids: bit-exact.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def chk():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from tools import gpu_check
    return gpu_check


def _all_below(res, bound, keys=None):
    for k, v in res.items():
        if keys is not None and not any(k.endswith(s) or s in k for s in keys):
            continue
        if isinstance(v, (list, tuple)):
            assert v[0] < bound, f"{k}: rel err {v[0]} >= {bound}"


def test_gemm_bias_act(chk):
    _all_below(chk.check_gemm_basic(), 4e-3)


def test_gemm_mn_major_operand(chk):
    _all_below(chk.check_gemm_kn(), 4e-3)


def test_gemm_gate_residual_and_hook_output(chk):
    _all_below(chk.check_gemm_gate(), 4e-3)


def test_qkv_rmsnorm_rope_epilogue(chk):
    _all_below(chk.check_qkv(), 4e-3)


def test_attention_shapes_and_ragged_tails(chk):
    res = chk.check_attention()
    assert len(res) >= 7
    _all_below(res, 6e-3)


def test_rowwise_kernels(chk):
    r = chk.check_rowwise()
    _all_below(r, 4e-3)
    assert r["rope_cos_maxabs"] <= 6e-8 and r["rope_sin_maxabs"] <= 6e-8  # at most 1 ulp; measured: bit-exact
    assert r["rope_bitexact_frac"] > 0.9999
    assert r["rope_compact_consistent"] is True
    assert r["euler_exact"] is True


def test_kd_loss_forward_backward_and_guard(chk):
    r = chk.check_kd()
    got, ref = r["loss"]
    assert abs(got - ref) / abs(ref) < 2e-3
    assert r["skipped_ref"] == [2] and r["valid"] == [1, 1, 0, 1]
    assert r["grad"][0] < 1e-2
    assert r["grad_skipped_layer_zero"] is True


def test_projector_front_end(chk):
    _all_below(chk.check_proj(), 4e-3)


def test_error_paths_fail_loudly():
    from x2i_b200 import ops
    from x2i_b200._lib import X2IError
    a = torch.randn(8, 24, device="cuda").bfloat16()
    w = torch.randn(30, 24, device="cuda").bfloat16()  # N % 32 != 0
    with pytest.raises(X2IError):
        ops.linear(a, w)
    with pytest.raises(X2IError):
        ops.linear(a.float(), w)
    with pytest.raises(X2IError):
        ops.attention(torch.randn(1, 1, 8, 64, device="cuda").bfloat16(), torch.randn(1, 1, 8, 64, device="cuda").bfloat16(),
                      torch.randn(1, 1, 8, 64, device="cuda").bfloat16())


def test_grouped_and_vae_entry_points_reject_bad_shapes():
    """The entry points added for the ControlNeXt stack / VAE fail loudly instead of reading out of bounds."""
    from x2i_b200 import ops
    from x2i_b200._lib import X2IError
    x = torch.randn(3, 16, 16, 64, device="cuda").bfloat16()
    w = torch.randn(2, 64, 9 * 64, device="cuda").bfloat16()
    b = torch.zeros(2, 64, device="cuda").bfloat16()
    with pytest.raises(X2IError):          # 3 images cannot be split over 2 weight sets
        ops.conv2d_nhwc(x, w, b, 3, 3, groups=2)
    with pytest.raises(X2IError):          # weight tensor does not match `groups`
        ops.conv2d_nhwc(x, w[0], b[0], 3, 3, groups=2)
    g = torch.ones(2, 64, device="cuda").bfloat16()
    with pytest.raises(X2IError):
        ops.groupnorm_nhwc(x, g, g, 8, 1e-6, param_sets=2)
    with pytest.raises(X2IError):          # 96 channels: not a power of two
        ops.groupnorm_nhwc(torch.randn(1, 8, 8, 96, device="cuda").bfloat16(), g[0, :64].repeat(2)[:96].contiguous(), g[0, :64].repeat(2)[:96].contiguous(), 4, 1e-6)
    with pytest.raises(X2IError):          # soft-max rows longer than the register-resident limit
        ops.softmax_rows(torch.zeros(2, 16388, device="cuda"))
    with pytest.raises(X2IError):          # fp32 GEMM needs N % 32 == 0
        ops.linear_f32(torch.randn(8, 64, device="cuda").bfloat16(), torch.randn(24, 64, device="cuda").bfloat16())


@pytest.mark.parametrize("B,C,S,H", [(1, 5, 128, 512), (2, 37, 256, 2048), (1, 29, 128, 3584), (1, 3, 384, 520)])
def test_layer_mixing_conv_on_the_tensor_pipe(chk, B, C, S, H):
    """x2i_proj_mix_ln_tc (banded-Toeplitz tcgen05 form of utils/proj.py's Conv2d(C -> 1, 5x5) + LayerNorm) against torch's fp32
    convolution on the same bf16 inputs and bf16 taps, and against the FP32-pipe stencil kernel it replaces for these shapes.
    H = 520: the last 56-column tile is partial and the first starts left of the plane (zero padding = TMA out-of-bounds fill)."""
    import torch.nn.functional as F
    from x2i_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    x = torch.randn(B, C, S, H, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 25, device="cuda", generator=g) * 0.1).bfloat16().float()
    gamma = (1 + 0.1 * torch.randn(H, device="cuda", generator=g)).float()
    beta = (0.1 * torch.randn(H, device="cuda", generator=g)).float()
    ref = F.conv2d(x.float(), w.view(1, C, 5, 5), torch.tensor([0.25], device="cuda"), padding=2)[:, 0]
    refn = F.layer_norm(ref, (H,), gamma, beta, 1e-6)
    from x2i_b200 import _lib
    assert _lib.lib().x2i_proj_mix_ln_tc_supported(B, C, S, H) == 1
    assert ops.proj_conv_tensor_cores
    y, xm = ops.proj_mix_ln_save(x, 0, w, 0.25, gamma, beta, 1e-6)
    ops.proj_conv_tensor_cores = False
    try:
        y0, xm0 = ops.proj_mix_ln_save(x, 0, w, 0.25, gamma, beta, 1e-6)
    finally:
        ops.proj_conv_tensor_cores = True
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())  # noqa: E731
    assert rel(xm, ref) < 4e-3 and rel(y, refn) < 4e-3          # one bf16 rounding of the output
    assert rel(y, y0) < 1e-3 and rel(xm, xm0) < 1e-3              # both kernels accumulate in fp32: they differ by summation order only
    assert torch.equal(ops.proj_mix_ln(x, 0, w, 0.25, gamma, beta, 1e-6), y)
