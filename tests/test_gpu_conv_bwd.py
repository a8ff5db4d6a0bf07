"""-m gpu: building blocks of the LightControl trainer (the trainer itself is not assembled yet): GroupNorm
backward and convolution input gradients on the sm_100a kernels against torch autograd in fp32 on the same bf16 inputs.
Tolerance: one bf16 rounding of the output (4e-3 relative) for dx, 1e-2 for the parameter sums."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import ops
    return ops


def _rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.mark.parametrize("C,G,act,hw", [(64, 2, 1, (40, 24)), (128, 4, 2, (33, 20)), (256, 8, 0, (16, 16)), (256, 8, 2, (64, 48))])
def test_groupnorm_backward(ops, C, G, act, hw):
    g = torch.Generator(device="cuda").manual_seed(C + act)
    x = (torch.randn(2, *hw, C, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(2, *hw, C, device="cuda", generator=g).bfloat16()
    ga = (torch.randn(C, device="cuda", generator=g) * 0.5 + 1).bfloat16()
    be = (torch.randn(C, device="cuda", generator=g) * 0.5).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = ga.float().requires_grad_(True), be.float().requires_grad_(True)
    y = F.group_norm(xr, G, gr, br, 1e-6)
    y = {0: y, 1: torch.relu(y), 2: F.silu(y)}[act]
    y.backward(dy.float().permute(0, 3, 1, 2))
    dx, dgamma, dbeta = ops.groupnorm_nhwc_bwd(x, dy, ga, be, G, 1e-6, act=act)
    assert _rel(dx, xr.grad.permute(0, 2, 3, 1)) < 4e-3
    assert _rel(dgamma, gr.grad) < 1e-2 and _rel(dbeta, br.grad) < 1e-2
    dx2, dg2, db2 = ops.groupnorm_nhwc_bwd(x, dy, ga, be, G, 1e-6, act=act, dgamma=dgamma.clone(), dbeta=dbeta.clone(), accumulate=True)
    assert torch.equal(dx, dx2)                                   # deterministic
    assert torch.allclose(dg2, 2 * dgamma, rtol=1e-6, atol=1e-6) and torch.allclose(db2, 2 * dbeta, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw", [(64, 64, 3, 1, 1, (24, 40)), (128, 256, 3, 1, 1, (16, 16)), (128, 256, 1, 1, 0, (16, 32)),
                                                        (128, 128, 3, 2, 1, (32, 48)), (256, 3072, 2, 2, 0, (16, 24))])
def test_conv_input_gradient(ops, cin, cout, k, stride, pad, hw):
    g = torch.Generator(device="cuda").manual_seed(cin + cout + k)
    x = torch.randn(2, cin, *hw, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.05).bfloat16()
    xr = x.float().requires_grad_(True)
    y = F.conv2d(xr, w.float(), None, stride=stride, padding=pad)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16()
    y.backward(dy.float())
    dx = ops.conv2d_nhwc_dgrad(dy.permute(0, 2, 3, 1).contiguous(), w, stride=stride, pad=pad)
    assert dx.shape == (2, *hw, cin)
    assert _rel(dx, xr.grad.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw", [(64, 64, 3, 1, 1, (24, 40)), (128, 256, 3, 1, 1, (16, 16)), (128, 256, 1, 1, 0, (16, 32)),
                                                        (128, 128, 3, 2, 1, (32, 48)), (256, 3072, 2, 2, 0, (16, 24)),
                                                        # output rows that tile into 64-pixel blocks: the implicit (im2col-free) kernel
                                                        (64, 128, 3, 1, 1, (8, 128)), (64, 64, 3, 1, 1, (16, 16)), (128, 128, 3, 2, 1, (32, 256)),
                                                        (128, 128, 3, 2, 1, (32, 32)), (256, 3072, 2, 2, 0, (16, 128)), (128, 64, 3, 1, 1, (4, 192))])
def test_conv_weight_and_bias_gradient(ops, cin, cout, k, stride, pad, hw):
    g = torch.Generator(device="cuda").manual_seed(cin + cout + k + 1)
    x = torch.randn(2, cin, *hw, device="cuda", generator=g).bfloat16()
    wr = (torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.05).requires_grad_(True)
    br = torch.zeros(cout, device="cuda").requires_grad_(True)
    y = F.conv2d(x.float(), wr, br, stride=stride, padding=pad)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16()
    y.backward(dy.float())
    dw, db = ops.conv2d_nhwc_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), k, k, stride=stride, pad=pad)
    assert _rel(ops.unpack_conv_weight_grad(dw, cin, k, k), wr.grad) < 4e-3
    assert _rel(db, br.grad) < 1e-3
    dw2, db2 = ops.conv2d_nhwc_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), k, k, stride=stride, pad=pad,
                                     dw=dw.clone(), db=db.clone(), accumulate=True)
    assert _rel(dw2, 2 * dw.float()) < 4e-3 and torch.allclose(db2, 2 * db, rtol=1e-5, atol=1e-4)
    ho, wo = y.shape[2], y.shape[3]
    implicit = (wo % 64 == 0) or (64 % wo == 0 and (ho * wo) % 64 == 0)
    from x2i_b200 import _lib
    assert bool(_lib.lib().x2i_conv2d_nhwc_wgrad_supported(hw[0], hw[1], cin, cout, k, k, stride, pad, pad)) == implicit
    if implicit:  # same result through the explicit im2col + GEMM form (one bf16 rounding per image there, one in total here)
        ops.implicit_conv_wgrad = False
        try:
            dw3, _ = ops.conv2d_nhwc_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), k, k, stride=stride, pad=pad)
        finally:
            ops.implicit_conv_wgrad = True
        assert _rel(dw3, dw.float()) < 6e-3


def test_conv_weight_gradient_asymmetric_padding(ops):
    """The VAE encoder's down-sampling convs: F.pad(x, (0, 1, 0, 1)) + 3x3 / stride 2 / no padding = pad 0, pad_end 1."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 128, 32, 128, device="cuda", generator=g).bfloat16()
    wr = (torch.randn(128, 128, 3, 3, device="cuda", generator=g) * 0.05).requires_grad_(True)
    y = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), wr, None, stride=2)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16()
    y.backward(dy.float())
    dw, _ = ops.conv2d_nhwc_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), 3, 3, stride=2, pad=0, pad_end=1)
    assert _rel(ops.unpack_conv_weight_grad(dw, 128, 3, 3), wr.grad) < 4e-3
