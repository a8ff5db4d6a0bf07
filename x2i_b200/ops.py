"""Tensor-level wrappers over the C ABI.  PyTorch is plumbing here (device memory + streams): every op below
enqueues hand-written sm_100a kernels from libx2i_b200.so on torch's current stream.  No fallbacks."""
import os

import torch

from . import _lib

BF16 = torch.bfloat16


def _p(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype=BF16):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.X2IError(f"{name} must be a CUDA tensor (x2i_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.X2IError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() > 0 and t.stride(-1) != 1:
        raise _lib.X2IError(f"{name} must be contiguous in its last dimension")


def _rows(t):
    """View [.., K] tensor as (rows, ld); requires a uniform row stride."""
    if t.dim() == 2:
        return t.shape[0], t.stride(0)
    if not t.is_contiguous():
        raise _lib.X2IError("expected a contiguous tensor or a 2-D row-strided view")
    return t.numel() // t.shape[-1], t.shape[-1]


def linear(x, weight, bias=None, act=0, out=None):
    """act(x @ weight.T + bias); act: 0 none, 1 gelu-tanh, 2 gelu-erf.  x: [..., K] -> [..., N]."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias")
    M, lda = _rows(x)
    N, K = weight.shape
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=BF16)
    _, ldc = _rows(out)
    _lib.call("x2i_gemm_bias_act", _p(x), lda, _p(weight), weight.stride(0), _p(bias), _p(out), ldc, M, N, K, act, _stream())
    return out


def linear_dual_gelu(x, weight, bias=None):
    """(y, gelu_erf(y)) with y = x @ weight.T + bias, one GEMM pass."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias")
    M, lda = _rows(x)
    N, K = weight.shape
    y = torch.empty(*x.shape[:-1], N, device=x.device, dtype=BF16)
    g = torch.empty_like(y)
    _lib.call("x2i_gemm_bias_dual", _p(x), lda, _p(weight), weight.stride(0), _p(bias), _p(y), N, _p(g), N, M, N, K, _stream())
    return y, g


def linear_gate_residual(x, weight, bias, gate, residual, rows_per_batch, out=None, aux=None):
    """residual + gate[b] * (x @ weight.T + bias); gate: [B, N] view (row-strided ok); aux receives the un-gated value."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias"); _chk(gate, "gate"); _chk(residual, "residual"); _chk(aux, "aux")
    M, lda = _rows(x)
    N, K = weight.shape
    if out is None:
        out = residual
    _, ldr = _rows(residual)
    _, ldc = _rows(out)
    ldaux = _rows(aux)[1] if aux is not None else 0
    _lib.call("x2i_gemm_gate_residual", _p(x), lda, _p(weight), weight.stride(0), _p(bias), _p(gate), gate.stride(0),
              rows_per_batch, _p(residual), ldr, _p(out), ldc, _p(aux), ldaux, M, N, K, _stream())
    return out


def qkv_rope(x, weight, bias, rms_q, rms_k, rope, q, k, v, heads, rows_per_batch, row_offset, eps=1e-6, mlp=None):
    """Fused QKV(+MLP) projection with RMSNorm + RoPE epilogue writing head-major q/k/v[B, heads, L, 128]."""
    for t, n in ((x, "x"), (weight, "weight"), (bias, "bias"), (rms_q, "rms_q"), (rms_k, "rms_k"), (q, "q"), (k, "k"), (v, "v"), (mlp, "mlp")):
        _chk(t, n)
    if rope is not None and (rope.dtype != torch.float32 or not rope.is_contiguous()):
        raise _lib.X2IError("rope must be the contiguous fp32 [L, 64, 2] table from rope_table()")
    M, lda = _rows(x)
    N, K = weight.shape
    L_total = q.shape[2]
    if not weight.is_contiguous() and weight.stride(1) != 1:
        raise _lib.X2IError("qkv_rope: weight rows must be contiguous")
    ldmlp = _rows(mlp)[1] if mlp is not None else 0
    _lib.call("x2i_gemm_qkv_rope", _p(x), lda, _p(weight), weight.stride(0), _p(bias), _p(rms_q), _p(rms_k), _p(rope), _p(q),
              _p(k), _p(v), _p(mlp), ldmlp, M, N, K, heads, rows_per_batch, row_offset, L_total, eps, _stream())


def desc_linear(x, weight, bias, out, act=0):
    """Descriptor of act(x @ weight.T + bias) -> out, for gemm_grouped()."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias"); _chk(out, "out")
    M, lda = _rows(x)
    N, K = weight.shape
    return _lib.GemmDesc(kind=_lib.GEMM_BIAS_ACT, act=act, M=M, N=N, K=K, A=_p(x), W=_p(weight), bias=_p(bias), lda=lda,
                         ldw=weight.stride(0), C=_p(out), ldc=_rows(out)[1])


def desc_gate_residual(x, weight, bias, gate, residual, rows_per_batch, aux=None, out=None):
    """Descriptor of out = residual + gate[b] * (x @ weight.T + bias) (out defaults to residual: in place); aux receives the
    un-gated value."""
    for t, n in ((x, "x"), (weight, "weight"), (bias, "bias"), (gate, "gate"), (residual, "residual"), (aux, "aux"), (out, "out")):
        _chk(t, n)
    M, lda = _rows(x)
    N, K = weight.shape
    ldr = _rows(residual)[1]
    if out is None:
        out = residual
    return _lib.GemmDesc(kind=_lib.GEMM_GATE_RESIDUAL, M=M, N=N, K=K, rows_per_batch=rows_per_batch, A=_p(x), W=_p(weight),
                         bias=_p(bias), lda=lda, ldw=weight.stride(0), C=_p(out), ldc=_rows(out)[1], gate=_p(gate),
                         residual=_p(residual), gate_stride=gate.stride(0), ldr=ldr, aux=_p(aux),
                         ldaux=_rows(aux)[1] if aux is not None else 0)


def desc_qkv_rope(x, weight, bias, rms_q, rms_k, rope, q, k, v, heads, rows_per_batch, row_offset, eps=1e-6, mlp=None):
    for t, n in ((x, "x"), (weight, "weight"), (bias, "bias"), (rms_q, "rms_q"), (rms_k, "rms_k"), (q, "q"), (k, "k"), (v, "v"), (mlp, "mlp")):
        _chk(t, n)
    M, lda = _rows(x)
    N, K = weight.shape
    return _lib.GemmDesc(kind=_lib.GEMM_QKV_ROPE, M=M, N=N, K=K, rows_per_batch=rows_per_batch, heads=heads, row_offset=row_offset,
                         L_total=q.shape[2], eps=eps, A=_p(x), W=_p(weight), bias=_p(bias), lda=lda, ldw=weight.stride(0),
                         rms_q=_p(rms_q), rms_k=_p(rms_k), rope=_p(rope), q=_p(q), k=_p(k), v=_p(v), mlp=_p(mlp),
                         ldmlp=_rows(mlp)[1] if mlp is not None else 0)


def gemm_grouped(*descs):
    """Run 1-2 GEMM descriptors of the same kind in ONE persistent CTA-pair launch (image + text stream of a block)."""
    arr = (_lib.GemmDesc * len(descs))(*descs)
    _lib.call("x2i_gemm_grouped", arr, len(descs), _stream())


def matmul_kn(a, b_kn, bias=None):
    """a[M,K] @ b_kn[K,N] with b N-contiguous (exercises the MN-major tcgen05 operand path used for V)."""
    _chk(a, "a"); _chk(b_kn, "b"); _chk(bias, "bias")
    M, K = a.shape
    N = b_kn.shape[1]
    out = torch.empty(M, N, device=a.device, dtype=BF16)
    _lib.call("x2i_gemm_kn", _p(a), a.stride(0), _p(b_kn), b_kn.stride(0), _p(bias), _p(out), N, M, N, K, _stream())
    return out


ATTN_EVENTS = None  # set to a list to collect (start, end) CUDA events around every attention launch (bench.py)


def attention(q, k, v, split=0, out0=None, out1=None):
    """softmax(q k^T / sqrt(128)) v for q,k,v [B, H, L, 128] -> token-major outputs.
    Rows t < split -> out0[B, split, H*128]; rows t >= split -> out1[B, L-split, >= H*128] (row-strided view ok)."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v")
    B, H, L, d = q.shape
    if d != 128 or not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise _lib.X2IError("attention: q,k,v must be contiguous [B,H,L,128]")
    if split > 0 and out0 is None:
        out0 = torch.empty(B, split, H * 128, device=q.device, dtype=BF16)
    if split < L and out1 is None:
        out1 = torch.empty(B, L - split, H * 128, device=q.device, dtype=BF16)
    _chk(out0, "out0"); _chk(out1, "out1")
    ld0 = out0.stride(-2) if out0 is not None else 0
    ld1 = out1.stride(-2) if out1 is not None else 0
    if ATTN_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("x2i_mmdit_attention", _p(q), _p(k), _p(v), _p(out0), ld0, split, _p(out1), ld1, B, H, L, _stream())
    if ATTN_EVENTS is not None:
        e1.record()
        ATTN_EVENTS.append((e0, e1))
    return out0, out1


def cross_attention(q, k, v, kv_len=None):
    """softmax(q k^T / sqrt(128) [+ key padding mask]) v; q [B,H,Lq,128], k,v [B,H,Lkv,128]; kv_len int32 [B] or None.
    Returns [B, Lq, H*128]."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v"); _chk(kv_len, "kv_len", torch.int32)
    B, H, Lq, d = q.shape
    Lkv = k.shape[2]
    if d != 128 or not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()) or v.shape != k.shape:
        raise _lib.X2IError("cross_attention: q [B,H,Lq,128], k,v [B,H,Lkv,128] contiguous")
    out = torch.empty(B, Lq, H * 128, device=q.device, dtype=BF16)
    _lib.call("x2i_cross_attention", _p(q), _p(k), _p(v), _p(kv_len), 0, 0, 0, _p(out), H * 128, B, H, Lq, Lkv, _stream())
    return out


def layernorm_affine(x, gamma, beta, eps=1e-6):
    _chk(x, "x"); _chk(gamma, "gamma"); _chk(beta, "beta")
    rows, ldx = _rows(x)
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.call("x2i_layernorm_affine", _p(x), ldx, _p(gamma), _p(beta), _p(y), x.shape[-1], rows, x.shape[-1], eps, _stream())
    return y


def add_pos2d(x, pos, tgt_sizes):
    """x [B,L,D] + per-image slice of the 2-D table pos [max_h,max_w,D] (zero beyond h_b*w_b); tgt_sizes int32 [B,2]."""
    _chk(x, "x"); _chk(pos, "pos"); _chk(tgt_sizes, "tgt_sizes", torch.int32)
    B, L, D = x.shape
    out = torch.empty_like(x)
    _lib.call("x2i_add_pos2d", _p(x.contiguous()), _p(pos.contiguous()), _p(tgt_sizes.contiguous()), _p(out), B, L, D,
              pos.shape[0], pos.shape[1], _stream())
    return out


class _NoRange:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NVTX_ON = os.environ.get("X2I_NVTX", "0") == "1"
_NO_RANGE = _NoRange()


def nvtx(name):
    """NVTX range around a phase of the path (`X2I_NVTX=1`; shows up in nsys / ncu --nvtx).  A no-op object otherwise, so the
    hot path pays one attribute lookup."""
    if not _NVTX_ON:
        return _NO_RANGE
    return torch.cuda.nvtx.range(name)


def ln_modulate2(x0, scale0, shift0, rpb0, out0, x1, scale1, shift1, rpb1, out1, eps=1e-6):
    """ln_modulate of two row segments (image + text stream of a double block) in ONE launch; outputs are written in place."""
    for t, n in ((x0, "x0"), (scale0, "scale0"), (shift0, "shift0"), (out0, "out0"), (x1, "x1"), (scale1, "scale1"), (shift1, "shift1"), (out1, "out1")):
        _chk(t, n)
    r0, ldx0 = _rows(x0)
    r1, ldx1 = _rows(x1)
    D = x0.shape[-1]
    if x1.shape[-1] != D or scale0.stride(0) != shift0.stride(0) or scale1.stride(0) != shift1.stride(0):
        raise _lib.X2IError("ln_modulate2: both segments share D; scale and shift of a segment share one row stride")
    _lib.call("x2i_ln_modulate2", _p(x0), ldx0, _p(scale0), _p(shift0), scale0.stride(0), _p(out0), _rows(out0)[1], r0, rpb0,
              _p(x1), ldx1, _p(scale1), _p(shift1), scale1.stride(0), _p(out1), _rows(out1)[1], r1, rpb1, D, eps, _stream())
    return out0, out1


def ln_modulate(x, scale, shift, rows_per_batch, eps=1e-6, out=None):
    """LayerNorm(x) * (1 + scale[b]) + shift[b]; scale/shift: [B, D] views sharing one row stride."""
    _chk(x, "x"); _chk(scale, "scale"); _chk(shift, "shift")
    rows, ldx = _rows(x)
    D = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _, ldy = _rows(out)
    if scale.stride(0) != shift.stride(0):
        raise _lib.X2IError("ln_modulate: scale and shift must share a row stride")
    _lib.call("x2i_ln_modulate", _p(x), ldx, _p(scale), _p(shift), scale.stride(0), _p(out), ldy, rows, D, rows_per_batch, eps, _stream())
    return out


def gate_residual_(x, y, gate, rows_per_batch):
    """x += gate[b] * y, in place; x, y: 2-D row-strided [rows, D]."""
    _chk(x, "x"); _chk(y, "y"); _chk(gate, "gate")
    rows, ldx = _rows(x)
    _, ldy = _rows(y)
    _lib.call("x2i_gate_residual", _p(x), ldx, _p(y), ldy, _p(gate), gate.stride(0), rows, x.shape[-1], rows_per_batch, _stream())
    return x


def skinny_linear(x, weight, bias=None, act_in=0, out=None, accumulate=False):
    """out[b] (+)= act_in(x[b]) @ weight.T + bias for a handful of rows (B <= 64); weights streamed once."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias")
    B, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(B, N, device=x.device, dtype=BF16)
    _lib.call("x2i_skinny_linear", _p(x), x.stride(0), _p(weight), weight.stride(0), _p(bias), _p(out), out.stride(0), B, N, K,
              act_in, 1 if accumulate else 0, _stream())
    return out


def timestep_sinusoid(t, dim=256):
    _chk(t, "t", torch.float32)
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=BF16)
    _lib.call("x2i_timestep_sinusoid", _p(t), _p(out), t.shape[0], dim, _stream())
    return out


def rope_table(ids, axes_dim=(16, 56, 56), theta=10000.0, full=True):
    """ids fp32 [L,3] -> (cos[L,128], sin[L,128], rope[L,64,2]) fp32."""
    _chk(ids, "ids", torch.float32)
    ids = ids.contiguous()
    L = ids.shape[0]
    D = sum(axes_dim)
    cos = torch.empty(L, D, device=ids.device, dtype=torch.float32) if full else None
    sin = torch.empty(L, D, device=ids.device, dtype=torch.float32) if full else None
    rope = torch.empty(L, D // 2, 2, device=ids.device, dtype=torch.float32)
    _lib.call("x2i_rope_table", _p(ids), L, axes_dim[0], axes_dim[1], axes_dim[2], float(theta), _p(cos), _p(sin), _p(rope), _stream())
    return cos, sin, rope


def euler_step_(x, v, dsigma):
    _chk(x, "x"); _chk(v, "v")
    if not (x.is_contiguous() and v.is_contiguous()):
        raise _lib.X2IError("euler_step_: contiguous tensors required")
    _lib.call("x2i_euler_step", _p(x), _p(v), float(dsigma), x.numel(), _stream())
    return x


proj_conv_tensor_cores = True  # tests flip this to compare the tcgen05 Toeplitz kernel with the FP32-pipe stencil kernel


def _proj_mix_ln_tc(x, w, conv_bias, gamma, beta, eps, want_xm):
    B, C, S, H = x.shape
    y = torch.empty(B, S, H, device=x.device, dtype=BF16)
    xm = torch.empty_like(y) if want_xm else None
    ws = _ws_f32("proj_conv_tc", _lib.lib().x2i_proj_mix_ln_tc_workspace_floats(B, C, S, H), x.device)
    _lib.call("x2i_proj_mix_ln_tc", _p(x), _p(w.contiguous()), float(conv_bias), _p(gamma.contiguous()), _p(beta.contiguous()), float(eps), _p(y),
              _p(xm), _p(ws), B, C, S, H, _stream())
    return y, xm


def proj_mix_ln(x, mode, w, conv_bias, gamma, beta, eps):
    """Projector front end: x bf16 [B,C,S,H] -> LayerNorm(mix(x)) bf16 [B,S,H].  Mode 0 (5x5 conv over the layers) runs on the tensor
    pipe (x2i_proj_mix_ln_tc) when the shape allows, otherwise -- and for the two mean modes -- on the streaming stencil kernel."""
    _chk(x, "x"); _chk(w, "w", torch.float32); _chk(gamma, "gamma", torch.float32); _chk(beta, "beta", torch.float32)
    x = x.contiguous()
    B, C, S, H = x.shape
    if mode == 0 and proj_conv_tensor_cores and _lib.lib().x2i_proj_mix_ln_tc_supported(B, C, S, H):
        return _proj_mix_ln_tc(x, w, conv_bias, gamma, beta, eps, False)[0]
    y = torch.empty(B, S, H, device=x.device, dtype=BF16)
    _lib.call("x2i_proj_mix_ln", _p(x), mode, _p(w), float(conv_bias), _p(gamma), _p(beta), float(eps), _p(y), B, C, S, H, _stream())
    return y


def mean_over_s(y):
    _chk(y, "y")
    B, S, N = y.shape
    out = torch.empty(B, N, device=y.device, dtype=BF16)
    _lib.call("x2i_mean_over_s", _p(y.contiguous()), _p(out), B, S, N, _stream())
    return out


def kd_loss_fwd(teacher, student, seg_row_start, seg_layer, n_layers, batch, temperature=3.0):
    """teacher/student: bf16 [rows, D]; seg_row_start int64 [n_seg+1], seg_layer int32 [n_seg] (device tensors).
    Returns (loss fp32 scalar tensor, layer_term[n_layers], valid[n_layers] int32)."""
    _chk(teacher, "teacher"); _chk(student, "student")
    _chk(seg_row_start, "seg_row_start", torch.int64); _chk(seg_layer, "seg_layer", torch.int32)
    rows, D = teacher.shape
    if not (teacher.is_contiguous() and student.is_contiguous()) or student.shape != teacher.shape:
        raise _lib.X2IError("kd_loss: teacher/student must be contiguous [rows, D] of equal shape")
    n_seg = seg_layer.numel()
    dev = teacher.device
    row_kl = torch.empty(rows, device=dev, dtype=torch.float32)
    seg_sum = torch.empty(n_seg, device=dev, dtype=torch.float64)
    layer_term = torch.empty(n_layers, device=dev, dtype=torch.float32)
    loss = torch.empty((), device=dev, dtype=torch.float32)
    valid = torch.empty(n_layers, device=dev, dtype=torch.int32)
    _lib.call("x2i_kd_loss_fwd", _p(teacher), _p(student), rows, D, float(temperature), _p(seg_row_start), _p(seg_layer),
              n_seg, n_layers, batch, _p(row_kl), _p(seg_sum), _p(layer_term), _p(loss), _p(valid), _stream())
    return loss, layer_term, valid


def kd_loss_bwd(teacher, student, seg_row_start, seg_layer, max_seg_rows, batch, valid, dloss, temperature=3.0):
    rows, D = teacher.shape
    row_scale = torch.empty(rows, device=teacher.device, dtype=torch.float32)
    grad = torch.empty_like(student)
    _chk(dloss, "dloss", torch.float32); _chk(valid, "valid", torch.int32)
    _lib.call("x2i_kd_loss_bwd", _p(teacher), _p(student), rows, D, float(temperature), _p(seg_row_start), _p(seg_layer),
              seg_layer.numel(), int(max_seg_rows), batch, _p(valid), _p(dloss), _p(row_scale), _p(grad), _stream())
    return grad


# ================================================================================================ backward / training
F32 = torch.float32


def linear_dgrad(dy, weight, pre=None, n_split=0, dact=1, addend=None, out=None):
    """dX = addend + (dy @ weight) * act'(pre) on columns >= n_split  (backward of y = x @ weight.T towards x).
    dy [.., Nout], weight [Nout, Kin] as stored by nn.Linear; pre [.., Kin - n_split] = pre-activation that produced x."""
    _chk(dy, "dy"); _chk(weight, "weight"); _chk(pre, "pre"); _chk(addend, "addend"); _chk(out, "out")
    M, lddy = _rows(dy)
    Nout, Kin = weight.shape
    if out is None:
        out = torch.empty(*dy.shape[:-1], Kin, device=dy.device, dtype=BF16)
    ldpre = _rows(pre)[1] if pre is not None else 0
    ldadd = _rows(addend)[1] if addend is not None else 0
    _lib.call("x2i_gemm_dgrad", _p(dy), lddy, _p(weight), weight.stride(0), _p(pre), ldpre, n_split, dact, _p(addend), ldadd,
              _p(out), _rows(out)[1], M, Nout, Kin, _stream())
    return out


def linear_wgrad(dy, x, out=None, accumulate=False):
    """dW[N, K] (+)= dy[M, N]^T @ x[M, K]  (weight gradient of y = x @ W.T)."""
    _chk(dy, "dy"); _chk(x, "x"); _chk(out, "out")
    M, lddy = _rows(dy)
    M2, ldx = _rows(x)
    if M != M2:
        raise _lib.X2IError("linear_wgrad: dy and x must have the same number of rows")
    N, K = dy.shape[-1], x.shape[-1]
    if out is None:
        out = torch.empty(N, K, device=dy.device, dtype=BF16)
        accumulate = False
    nws = _lib.lib().x2i_gemm_wgrad_workspace_floats(M, N, K) if dy.is_cuda else 0
    if nws > 0:  # few output tiles, long contraction: split-K with a deterministic reduction
        ws = _ws_f32("wgrad_splitk", nws, dy.device)
        _lib.call("x2i_gemm_wgrad_splitk", _p(dy), lddy, _p(x), ldx, _p(out), out.stride(0), M, N, K, 1 if accumulate else 0, _p(ws), nws,
                  _stream())
    else:
        _lib.call("x2i_gemm_wgrad", _p(dy), lddy, _p(x), ldx, _p(out), out.stride(0), M, N, K, 1 if accumulate else 0, _stream())
    return out


def linear_act_save(x, weight, bias, act, pre_out=None, act_out=None):
    """(pre, act(pre)) with pre = x @ weight.T + bias; act 1 gelu-tanh, 2 gelu-erf."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(bias, "bias"); _chk(pre_out, "pre_out"); _chk(act_out, "act_out")
    M, lda = _rows(x)
    N, K = weight.shape
    if pre_out is None:
        pre_out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=BF16)
    if act_out is None:
        act_out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=BF16)
    _lib.call("x2i_gemm_bias_act_save", _p(x), lda, _p(weight), weight.stride(0), _p(bias), _p(pre_out), _rows(pre_out)[1],
              _p(act_out), _rows(act_out)[1], M, N, K, act, _stream())
    return pre_out, act_out


def desc_linear_act_save(x, weight, bias, pre_out, act_out, act=1):
    d = desc_linear(x, weight, bias, pre_out, act=0)
    _chk(act_out, "act_out")
    d.aux = _p(act_out); d.ldaux = _rows(act_out)[1]; d.aux_act = act
    return d


def desc_qkv_rope_save(x, weight, bias, rms_q, rms_k, rope, q, k, v, heads, rows_per_batch, row_offset, qk_pre, eps=1e-6, mlp=None,
                       mlp_pre=None):
    d = desc_qkv_rope(x, weight, bias, rms_q, rms_k, rope, q, k, v, heads, rows_per_batch, row_offset, eps, mlp)
    _chk(qk_pre, "qk_pre"); _chk(mlp_pre, "mlp_pre")
    d.qk_pre = _p(qk_pre); d.ldqk = _rows(qk_pre)[1]
    if mlp_pre is not None:
        d.mlp_pre = _p(mlp_pre); d.ldmlp_pre = _rows(mlp_pre)[1]
    return d


def lse_pad(L):
    return (L + 127) // 128 * 128


def attention_lse(q, k, v, split=0, out0=None, out1=None, lse=None):
    """attention() that also returns the log2-domain row log-sum-exp [B, H, Lpad] fp32 for the backward."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v")
    B, H, L, d = q.shape
    if d != 128 or not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise _lib.X2IError("attention: q,k,v must be contiguous [B,H,L,128]")
    if split > 0 and out0 is None:
        out0 = torch.empty(B, split, H * 128, device=q.device, dtype=BF16)
    if split < L and out1 is None:
        out1 = torch.empty(B, L - split, H * 128, device=q.device, dtype=BF16)
    if lse is None:
        lse = torch.empty(B, H, lse_pad(L), device=q.device, dtype=F32)
    _chk(out0, "out0"); _chk(out1, "out1"); _chk(lse, "lse", F32)
    ld0 = out0.stride(-2) if out0 is not None else 0
    ld1 = out1.stride(-2) if out1 is not None else 0
    _lib.call("x2i_mmdit_attention_lse", _p(q), _p(k), _p(v), _p(out0), ld0, split, _p(out1), ld1, _p(lse), B, H, L, _stream())
    return out0, out1, lse


def attention_bwd_prep(do0, do1, o0, o1, B, H, L, split, add0=None, add1=None, do_hm=None, delta=None):
    """token-major dO (+addend) and O -> head-major dO [B,H,L,128] and delta [B,H,Lpad]; *0 = tokens < split, *1 = the rest."""
    for t, n in ((do0, "do0"), (do1, "do1"), (o0, "o0"), (o1, "o1"), (add0, "add0"), (add1, "add1")):
        _chk(t, n)
    dev = (do1 if do1 is not None else do0).device
    if do_hm is None:
        do_hm = torch.empty(B, H, L, 128, device=dev, dtype=BF16)
    if delta is None:
        delta = torch.empty(B, H, lse_pad(L), device=dev, dtype=F32)
    ld = lambda t: t.stride(-2) if t is not None else 0  # noqa: E731
    _lib.call("x2i_attention_bwd_prep", _p(do0), ld(do0), _p(do1), ld(do1), _p(o0), ld(o0), _p(o1), ld(o1), _p(add0), ld(add0),
              _p(add1), ld(add1), _p(do_hm), _p(delta), B, H, L, split, _stream())
    return do_hm, delta


def attention_bwd(q, k, v, do_hm, lse, delta, dq=None, dk=None, dv=None):
    """(dq, dk, dv) head-major for softmax(q k^T / sqrt(128)) v."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (do_hm, "do_hm")):
        _chk(t, n)
        if not t.is_contiguous():
            raise _lib.X2IError(f"attention_bwd: {n} must be contiguous [B,H,L,128]")
    _chk(lse, "lse", F32); _chk(delta, "delta", F32)
    B, H, L, _ = q.shape
    if lse.shape != (B, H, lse_pad(L)) or delta.shape != lse.shape or not (lse.is_contiguous() and delta.is_contiguous()):
        raise _lib.X2IError("attention_bwd: lse / delta must be contiguous [B, H, Lpad] fp32")
    dq = torch.empty_like(q) if dq is None else dq
    dk = torch.empty_like(q) if dk is None else dk
    dv = torch.empty_like(q) if dv is None else dv
    _lib.call("x2i_mmdit_attention_bwd", _p(q), _p(k), _p(v), _p(do_hm), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), B, H, L, _stream())
    return dq, dk, dv


def qk_norm_rope_bwd(dq, dk, dv, qk_pre, rms_q, rms_k, rope, out, rows_per_batch, row_offset, eps=1e-6):
    """head-major dq,dk,dv -> out[M, >=3D] = [dq_pre | dk_pre | dv] token-major for the stream's tokens."""
    for t, n in ((dq, "dq"), (dk, "dk"), (dv, "dv"), (qk_pre, "qk_pre"), (rms_q, "rms_q"), (rms_k, "rms_k"), (out, "out")):
        _chk(t, n)
    B, H, L_total, _ = dq.shape
    M = _rows(qk_pre)[0]
    _lib.call("x2i_qk_norm_rope_bwd", _p(dq), _p(dk), _p(dv), _p(qk_pre), _rows(qk_pre)[1], _p(rms_q), _p(rms_k), _p(rope), _p(out),
              _rows(out)[1], M, H, rows_per_batch, row_offset, L_total, eps, _stream())
    return out


def gate_bwd(dx, gate, rows_per_batch, addend=None, out=None):
    _chk(dx, "dx"); _chk(gate, "gate"); _chk(addend, "addend"); _chk(out, "out")
    rows, lddx = _rows(dx)
    if out is None:
        out = torch.empty(dx.shape, device=dx.device, dtype=BF16)
    _lib.call("x2i_gate_bwd", _p(dx), lddx, _p(gate), gate.stride(0), _p(addend), _rows(addend)[1] if addend is not None else 0,
              _p(out), _rows(out)[1], rows, dx.shape[-1], rows_per_batch, _stream())
    return out


def ln_modulate_bwd(dn, x, scale, rows_per_batch, dres=None, out=None, stats=None, eps=1e-6, affine=False):
    """dx = dres + d/dx [LN(x) * (1 + scale) + shift] . dn   (affine: LN(x) * scale + shift); stats float2 [rows] optional."""
    _chk(dn, "dn"); _chk(x, "x"); _chk(scale, "scale"); _chk(dres, "dres"); _chk(out, "out"); _chk(stats, "stats", F32)
    rows, lddn = _rows(dn)
    if out is None:
        out = torch.empty(dn.shape, device=dn.device, dtype=BF16)
    mod_stride = 0 if affine else scale.stride(0)
    _lib.call("x2i_ln_modulate_bwd", _p(dn), lddn, _p(x), _rows(x)[1], _p(scale), mod_stride, _p(dres),
              _rows(dres)[1] if dres is not None else 0, _p(out), _rows(out)[1], _p(stats), rows, dn.shape[-1], rows_per_batch,
              eps, 1 if affine else 0, _stream())
    return out


_COLSUM_WS = {}


def _ws_f32(key, n, device):
    """Scratch buffer of an op, one per (op, device, CUDA stream): two streams (e.g. the reference's background data-loader
    thread doing GPU work, core/data/dataloader.py:100-123) never share a workspace; on one stream the launches are ordered."""
    k = (key, str(device), _stream())
    t = _COLSUM_WS.get(k)
    if t is None or t.numel() < n:
        t = torch.empty(int(n), device=device, dtype=F32)
        _COLSUM_WS[k] = t
    return t


def colsum(a, nbatch, rows_per_batch, out0=None, b=None, out1=None, stats=None, accumulate=False):
    """out0[nb, D] (+)= sum_t a;  out1[nb, D] (+)= sum_t a * (b, or (b - mean) * rstd with stats).  fp32 outputs (row-strided ok)."""
    _chk(a, "a"); _chk(b, "b"); _chk(out0, "out0", F32); _chk(out1, "out1", F32); _chk(stats, "stats", F32)
    D = a.shape[-1]
    ws = _ws_f32("colsum", _lib.lib().x2i_colsum_workspace_floats(nbatch, rows_per_batch, D), a.device)
    _lib.call("x2i_colsum", _p(a), _rows(a)[1], _p(b), _rows(b)[1] if b is not None else 0, _p(stats), _p(out0),
              out0.stride(0) if out0 is not None else 0, _p(out1), out1.stride(0) if out1 is not None else 0, _p(ws), nbatch,
              rows_per_batch, D, 1 if accumulate else 0, _stream())


def skinny_linear_t(g, weight, pre=None, dact=0, out=None, accumulate=False):
    """out[b, k] (+)= act'(pre[b, k]) * sum_n g[b, n] * weight[n, k]; g, out fp32; dact 1 = SiLU'."""
    _chk(g, "g", F32); _chk(weight, "weight"); _chk(pre, "pre"); _chk(out, "out", F32)
    B, N = g.shape
    K = weight.shape[1]
    if out is None:
        out = torch.empty(B, K, device=g.device, dtype=F32)
        accumulate = False
    ws = _ws_f32("skinny_t", _lib.lib().x2i_skinny_linear_t_workspace_floats(N, K), g.device)
    _lib.call("x2i_skinny_linear_t", _p(g), g.stride(0), _p(weight), weight.stride(0), _p(pre), pre.stride(0) if pre is not None else 0,
              _p(out), out.stride(0), _p(ws), B, N, K, dact, 1 if accumulate else 0, _stream())
    return out


def f32_to_bf16(x):
    _chk(x, "x", F32)
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.call("x2i_f32_to_bf16", _p(x.contiguous()), _p(out), x.numel(), _stream())
    return out


def proj_mix_ln_save(x, mode, w, conv_bias, gamma, beta, eps):
    """proj_mix_ln that also returns the pre-LayerNorm mixed plane xm [B,S,H] (bf16) for the backward."""
    _chk(x, "x"); _chk(w, "w", F32); _chk(gamma, "gamma", F32); _chk(beta, "beta", F32)
    x = x.contiguous()
    B, C, S, H = x.shape
    if mode == 0 and proj_conv_tensor_cores and _lib.lib().x2i_proj_mix_ln_tc_supported(B, C, S, H):
        return _proj_mix_ln_tc(x, w, conv_bias, gamma, beta, eps, True)
    y = torch.empty(B, S, H, device=x.device, dtype=BF16)
    xm = torch.empty_like(y)
    _lib.call("x2i_proj_mix_ln_save", _p(x), mode, _p(w), float(conv_bias), _p(gamma), _p(beta), float(eps), _p(y), _p(xm), B, C, S, H,
              _stream())
    return y, xm


def mean_over_s_bwd(dpooled, S):
    _chk(dpooled, "dpooled")
    B, N = dpooled.shape
    dy = torch.empty(B, S, N, device=dpooled.device, dtype=BF16)
    _lib.call("x2i_mean_over_s_bwd", _p(dpooled.contiguous()), _p(dy), B, S, N, _stream())
    return dy


def proj_mix_wgrad(x, g, mode):
    """fp32 gradient of the conv weight [C,5,5] (mode 0) or of cha_scale [C] (mode 1); x bf16 [B,C,S,H], g bf16 [B,S,H]."""
    _chk(x, "x"); _chk(g, "g")
    B, C, S, H = x.shape
    dw = torch.empty((C, 5, 5) if mode == 0 else (C,), device=x.device, dtype=F32)
    ws = _ws_f32("proj_mix_wgrad", _lib.lib().x2i_proj_mix_wgrad_workspace_floats(B, C, S), x.device)
    _lib.call("x2i_proj_mix_wgrad", _p(x.contiguous()), _p(g.contiguous()), mode, _p(dw), _p(ws), B, C, S, H, _stream())
    return dw


# ================================================================================================ ControlNeXt (LightControl)
def conv2d_nhwc(x, w_packed, bias, kh, kw, stride=1, pad=1, rowvec=None, residual=None, relu=False, out=None, groups=1, pad_end=None):
    """NHWC implicit-GEMM conv: x [N,H,W,Cin], w_packed [Cout, kh*kw*Cin] (from pack_conv_weight), bias [Cout];
    out = relu?(conv + bias + rowvec[n]) + residual, [N,Ho,Wo,Cout].
    groups > 1: w_packed [groups, Cout, kh*kw*Cin], bias [groups, Cout]; image n uses weight set n // (N // groups).
    pad_end: bottom/right zero padding when it differs from the top/left `pad` (VAE encoder down-sampling: pad=0, pad_end=1)."""
    for t, n in ((x, "x"), (w_packed, "w"), (bias, "bias"), (rowvec, "rowvec"), (residual, "residual"), (out, "out")):
        _chk(t, n)
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[-2]
    ok = w_packed.shape[-1] == kh * kw * Cin and x.is_contiguous() and w_packed.is_contiguous()
    ok = ok and (w_packed.dim() == 2 if groups == 1 else (w_packed.dim() == 3 and w_packed.shape[0] == groups and bias.is_contiguous()))
    if not ok:
        raise _lib.X2IError("conv2d_nhwc: x must be contiguous NHWC and w packed [(groups,) Cout, kh*kw*Cin]")
    pad_end = pad if pad_end is None else pad_end
    Ho, Wo = (H + pad + pad_end - kh) // stride + 1, (W + pad + pad_end - kw) // stride + 1
    if out is None:
        out = torch.empty(N, Ho, Wo, Cout, device=x.device, dtype=BF16)
    if residual is not None and (residual.numel() != out.numel() or not residual.is_contiguous()):
        raise _lib.X2IError("conv2d_nhwc: residual must be contiguous with the output's shape")
    _lib.call("x2i_conv2d_nhwc_grouped", _p(x), _p(w_packed), _p(bias), _p(rowvec), rowvec.stride(0) if rowvec is not None else 0,
              _p(residual), _p(out), N, H, W, Cin, Cout, kh, kw, stride, pad, pad_end, 1 if relu else 0, groups, _stream())
    return out


def pack_conv_weight(w):
    """[Cout, Cin, KH, KW] -> bf16 [Cout, KH*KW*Cin] (tap-major, channels innermost): the B operand of conv2d_nhwc."""
    return w.detach().permute(0, 2, 3, 1).contiguous().view(w.shape[0], -1).to(BF16).contiguous()


def conv_first(x_nchw, w, bias):
    """Conv2d(3->64, 3x3, s2, p1): x bf16 [N,3,H,W] -> NHWC bf16 [N,H/2,W/2,64]; w fp32 [64,3,3,3], bias fp32 [64].
    Grouped form: w [G,64,3,3,3], bias [G,64] -> [G*N, H/2, W/2, 64] (G stems on the same input)."""
    _chk(x_nchw, "x"); _chk(w, "w", F32); _chk(bias, "bias", F32)
    N, C, H, W = x_nchw.shape
    groups = w.shape[0] if w.dim() == 5 else 1
    if C != 3 or tuple(w.shape[-4:]) != (64, 3, 3, 3) or bias.numel() != groups * 64:
        raise _lib.X2IError("conv_first: x [N,3,H,W], w [(G,)64,3,3,3], bias [(G,)64]")
    out = torch.empty(groups * N, H // 2, W // 2, 64, device=x_nchw.device, dtype=BF16)
    _lib.call("x2i_conv_first_grouped", _p(x_nchw.contiguous()), _p(w.contiguous()), _p(bias.contiguous()), _p(out), N, H, W, groups,
              _stream())
    return out


def groupnorm_nhwc(x, gamma, beta, groups, eps, act=0, residual=None, out=None, param_sets=1):
    """act(GroupNorm(x)) + residual on NHWC bf16; act 0 none, 1 relu, 2 silu.
    param_sets > 1: gamma / beta [param_sets, C]; image n uses set n // (N // param_sets)."""
    for t, n in ((x, "x"), (gamma, "gamma"), (beta, "beta"), (residual, "residual"), (out, "out")):
        _chk(t, n)
    N, H, W, C = x.shape
    if gamma.numel() != param_sets * C or beta.numel() != param_sets * C or not gamma.is_contiguous() or not beta.is_contiguous():
        raise _lib.X2IError("groupnorm_nhwc: gamma / beta must be contiguous [(param_sets,) C]")
    if out is None:
        out = torch.empty_like(x)
    ws = _ws_f32("groupnorm", _lib.lib().x2i_groupnorm_workspace_floats(N, H * W, groups), x.device)
    _lib.call("x2i_groupnorm_nhwc_grouped", _p(x), _p(gamma), _p(beta), _p(residual), _p(out), _p(ws), N, H * W, C, groups, float(eps), act,
              param_sets, _stream())
    return out


# ================================================================================================ VAE decoder
def linear_f32(x, weight, alpha=1.0, out=None):
    """fp32 out[M, N] = alpha * x[M, K] @ weight[N, K]^T (attention scores that must not be rounded to bf16)."""
    _chk(x, "x"); _chk(weight, "weight")
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=F32)
    _chk(out, "out", F32)
    _lib.call("x2i_gemm_f32", _p(x), x.stride(0), _p(weight), weight.stride(0), _p(out), out.stride(0), M, N, K, float(alpha), _stream())
    return out


def softmax_rows(s, out=None):
    """bf16 softmax over the last dim of fp32 scores [R, C]."""
    _chk(s, "s", F32)
    R, C = s.shape
    if out is None:
        out = torch.empty(R, C, device=s.device, dtype=BF16)
    _chk(out, "out")
    _lib.call("x2i_softmax_rows", _p(s), s.stride(0), _p(out), out.stride(0), R, C, _stream())
    return out


def softmax_rows_bias(s, bias, out=None):
    """P (bf16) = softmax(s + bias) row-wise; s fp32 [rows, cols], bias fp32 [bias_rows, cols] (row r uses bias row r % bias_rows)."""
    _chk(s, "s", F32); _chk(bias, "bias", F32)
    rows, cols = s.shape
    if out is None:
        out = torch.empty(rows, cols, device=s.device, dtype=BF16)
    _chk(out, "out")
    _lib.call("x2i_softmax_rows_bias", _p(s), s.stride(0), _p(bias), bias.stride(0), bias.shape[0], _p(out), out.stride(0), rows, cols, _stream())
    return out


def upsample2x_nhwc(x, out=None):
    """Nearest 2x upsampling of NHWC bf16."""
    _chk(x, "x")
    N, H, W, C = x.shape
    if not x.is_contiguous():
        raise _lib.X2IError("upsample2x_nhwc: x must be contiguous NHWC")
    if out is None:
        out = torch.empty(N, 2 * H, 2 * W, C, device=x.device, dtype=BF16)
    _lib.call("x2i_upsample2x_nhwc", _p(x), _p(out), N, H, W, C, _stream())
    return out


# ================================================================================================ LightControl trainer building blocks
def groupnorm_nhwc_bwd(x, dy, gamma, beta, groups, eps, act=0, dgamma=None, dbeta=None, accumulate=False):
    """Backward of act(GroupNorm(x)) on NHWC bf16: returns (dx bf16, dgamma fp32 [C], dbeta fp32 [C]).  act 0 none, 1 relu, 2 silu.
    With accumulate=True the parameter gradients are added into the given fp32 buffers (several images / nets per parameter)."""
    for t, n in ((x, "x"), (dy, "dy"), (gamma, "gamma"), (beta, "beta")):
        _chk(t, n)
    N, H, W, C = x.shape
    if dy.shape != x.shape or not x.is_contiguous() or not dy.is_contiguous() or gamma.numel() != C or beta.numel() != C:
        raise _lib.X2IError("groupnorm_nhwc_bwd: x, dy contiguous NHWC of equal shape; gamma, beta [C]")
    dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(C, device=x.device, dtype=F32)
        dbeta = torch.empty(C, device=x.device, dtype=F32)
        accumulate = False
    _chk(dgamma, "dgamma", F32); _chk(dbeta, "dbeta", F32)
    ws = _ws_f32("groupnorm_bwd", _lib.lib().x2i_groupnorm_bwd_workspace_floats(N, H * W, C, groups), x.device)
    _lib.call("x2i_groupnorm_nhwc_bwd", _p(x), _p(dy), _p(gamma.contiguous()), _p(beta.contiguous()), _p(dx), _p(dgamma), _p(dbeta), _p(ws),
              N, H * W, C, groups, float(eps), act, 1 if accumulate else 0, _stream())
    return dx, dgamma, dbeta


def pack_conv_weight_dgrad(w):
    """[Cout, Cin, KH, KW] -> the packed weight of the convolution that computes the INPUT gradient: channels swapped, taps
    flipped, i.e. pack_conv_weight(w.transpose(0, 1).flip(2, 3)) -> bf16 [Cin, KH*KW*Cout]."""
    return pack_conv_weight(w.detach().transpose(0, 1).flip(2, 3))


def conv2d_nhwc_dgrad(dy, w, stride=1, pad=1):
    """Input gradient of conv2d_nhwc for w in PyTorch layout [Cout, Cin, KH, KW]: dy [N, Ho, Wo, Cout] -> dx [N, H, W, Cin].
      stride 1: the forward implicit-GEMM kernel on the flipped / transposed weights with padding k - 1 - pad;
      stride 2 (3x3, pad 1, even H, W): the same on dy with zeros inserted between the pixels (Z[2o] = dy[o]);
      2x2 / stride 2 / pad 0 (non-overlapping patches): one dgrad GEMM on the forward-packed weight + depth-to-space."""
    _chk(dy, "dy"); _chk(w, "w")
    Cout, Cin, kh, kw = w.shape
    N, Ho, Wo, C = dy.shape
    if C != Cout:
        raise _lib.X2IError("conv2d_nhwc_dgrad: dy channels must equal w.shape[0]")
    if stride == 1 and kh == kw and 0 <= kh - 1 - pad <= 1:
        return conv2d_nhwc(dy.contiguous(), pack_conv_weight_dgrad(w), None, kh, kw, stride=1, pad=kh - 1 - pad)
    if stride == 2 and kh == 3 and kw == 3 and pad == 1:
        z = torch.zeros(N, 2 * Ho, 2 * Wo, Cout, device=dy.device, dtype=BF16)
        z[:, ::2, ::2] = dy
        return conv2d_nhwc(z, pack_conv_weight_dgrad(w), None, 3, 3, stride=1, pad=1)
    if stride == 2 and kh == 2 and kw == 2 and pad == 0:
        cols = linear_dgrad(dy.reshape(N * Ho * Wo, Cout), pack_conv_weight(w))           # [P, (ky, kx, Cin)]
        return cols.view(N, Ho, Wo, 2, 2, Cin).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * Ho, 2 * Wo, Cin).contiguous()
    raise _lib.X2IError("conv2d_nhwc_dgrad: supported forms are stride 1 (k - 1 - pad in {0, 1}), 3x3/s2/p1 and 2x2/s2/p0")


implicit_conv_wgrad = True  # tests flip this to compare the implicit kernel with the explicit im2col + GEMM form


def conv2d_nhwc_wgrad(x, dy, kh, kw, stride=1, pad=1, pad_end=None, dw=None, db=None, accumulate=False):
    """Weight and bias gradient of conv2d_nhwc: x [N, H, W, Cin], dy [N, Ho, Wo, Cout] -> (dW bf16 in the PACKED layout
    [Cout, kh*kw*Cin] of pack_conv_weight, db fp32 [Cout]).  Implicit form (x2i_conv2d_nhwc_wgrad: tap-shifted TMA boxes of x as the
    GEMM's B operand, all images in one launch) when the output rows tile into 64-pixel blocks; otherwise the explicit form: one image
    at a time, im2col (x2i_im2col_nhwc) + the MN-major wgrad GEMM accumulating over the images.  db = column sums of dy.
    unpack_conv_weight_grad() gives PyTorch's layout."""
    _chk(x, "x"); _chk(dy, "dy"); _chk(dw, "dw"); _chk(db, "db", F32)
    N, H, W, Cin = x.shape
    _, Ho, Wo, Cout = dy.shape
    pad_end = pad if pad_end is None else pad_end
    if (Ho, Wo) != ((H + pad + pad_end - kh) // stride + 1, (W + pad + pad_end - kw) // stride + 1) or dy.shape[0] != N:
        raise _lib.X2IError("conv2d_nhwc_wgrad: dy does not match the convolution's output shape")
    if not x.is_contiguous() or not dy.is_contiguous():
        raise _lib.X2IError("conv2d_nhwc_wgrad: x and dy must be contiguous NHWC")
    if dw is None:
        dw = torch.empty(Cout, kh * kw * Cin, device=x.device, dtype=BF16)
        db = torch.empty(Cout, device=x.device, dtype=F32)
        accumulate = False
    L = _lib.lib()
    if implicit_conv_wgrad and L.x2i_conv2d_nhwc_wgrad_supported(H, W, Cin, Cout, kh, kw, stride, pad, pad_end):
        # implicit form: the GEMM's B operand is x itself, loaded as tap-shifted boxes -- one launch for all images, no im2col buffer
        nws = L.x2i_conv2d_nhwc_wgrad_workspace_floats(N, H, W, Cin, Cout, kh, kw, stride, pad, pad_end)
        ws = _ws_f32("conv_wgrad_splitk", nws, x.device) if nws > 0 else None
        _lib.call("x2i_conv2d_nhwc_wgrad", _p(x), _p(dy), _p(dw), _p(ws), nws, N, H, W, Cin, Cout, kh, kw, stride, pad, pad_end,
                  1 if accumulate else 0, _stream())
    else:  # explicit form (output rows that do not tile into 64-pixel blocks)
        cols = torch.empty(Ho * Wo, kh * kw * Cin, device=x.device, dtype=BF16)
        for n in range(N):
            _lib.call("x2i_im2col_nhwc", _p(x[n]), _p(cols), 1, H, W, Cin, kh, kw, stride, pad, pad_end, _stream())
            linear_wgrad(dy[n].reshape(Ho * Wo, Cout), cols, out=dw, accumulate=accumulate or n > 0)
    colsum(dy.reshape(N * Ho * Wo, Cout), 1, N * Ho * Wo, out0=db.view(1, Cout), accumulate=accumulate)
    return dw, db


def unpack_conv_weight_grad(dw_packed, cin, kh, kw):
    """[Cout, kh*kw*Cin] (tap-major, channels innermost) -> PyTorch's [Cout, Cin, kh, kw]."""
    return dw_packed.view(dw_packed.shape[0], kh, kw, cin).permute(0, 3, 1, 2).contiguous()


def relu_bwd(dy, y):
    """dy masked by y > 0 (backward of a ReLU fused into a conv epilogue)."""
    _chk(dy, "dy"); _chk(y, "y")
    if dy.shape != y.shape or not dy.is_contiguous() or not y.is_contiguous():
        raise _lib.X2IError("relu_bwd: dy and y must be contiguous and of equal shape")
    out = torch.empty_like(dy)
    _lib.call("x2i_relu_bwd", _p(dy), _p(y), _p(out), dy.numel(), _stream())
    return out


def silu_rows(x):
    """silu(x) on a small bf16 tensor (time-embedding rows)."""
    _chk(x, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.call("x2i_silu", _p(x), _p(out), x.numel(), _stream())
    return out


# ================================================================================================ MLLM prefill
def _bstrided(t, name):
    """[B, S, D] view with unit last stride and one row stride inside a batch -> (ld, batch_stride)."""
    _chk(t, name)
    if t.dim() != 3:
        raise _lib.X2IError(f"{name} must be a [B, S, D] view")
    return t.stride(1), t.stride(0)


def gather_rows(ids, table, out):
    """out[b, s, :] = table[ids[b, s], :] (nn.Embedding).  ids int64 [B, S]; out: [B, S, D] view (e.g. layer slot 0 of [B, C, S, D])."""
    _chk(ids, "ids", torch.int64); _chk(table, "table")
    B, S = ids.shape
    ldo, bso = _bstrided(out, "out")
    _lib.call("x2i_gather_rows", _p(ids.contiguous()), _p(table), table.stride(0), table.shape[0], _p(out), ldo, bso, B * S, S,
              table.shape[1], _stream())
    return out


def rmsnorm(x, weight, eps=1e-6, out=None):
    """Qwen2RMSNorm on a [B, S, D] view -> [B, S, D] (dense unless `out` is given)."""
    _chk(weight, "weight")
    B, S, D = x.shape
    ldx, bsx = _bstrided(x, "x")
    if out is None:
        out = torch.empty(B, S, D, device=x.device, dtype=BF16)
    ldy, bsy = _bstrided(out, "out")
    _lib.call("x2i_rmsnorm", _p(x), ldx, bsx, _p(weight), _p(out), ldy, bsy, B * S, S, D, float(eps), _stream())
    return out


def rope_half_split(qkv, pos, inv_freq, heads, heads_kv, q=None, k=None, v=None):
    """Fused QKV rows [B, S, (heads + 2 heads_kv) * 128] -> rotate-half RoPE'd head-major q [B,heads,S,128], k, v [B,heads_kv,S,128]."""
    _chk(qkv, "qkv"); _chk(pos, "pos", torch.int32); _chk(inv_freq, "inv_freq", torch.float32)
    B, S, W = qkv.shape
    if not qkv.is_contiguous() or W != (heads + 2 * heads_kv) * 128 or pos.numel() != B * S or inv_freq.numel() != 64:
        raise _lib.X2IError("rope_half_split: qkv [B,S,(H+2Hkv)*128] contiguous, pos int32 [B,S], inv_freq fp32 [64]")
    dev = qkv.device
    q = q if q is not None else torch.empty(B, heads, S, 128, device=dev, dtype=BF16)
    k = k if k is not None else torch.empty(B, heads_kv, S, 128, device=dev, dtype=BF16)
    v = v if v is not None else torch.empty(B, heads_kv, S, 128, device=dev, dtype=BF16)
    _lib.call("x2i_rope_half_split", _p(qkv), W, _p(pos.contiguous()), _p(inv_freq.contiguous()), _p(q), _p(k), _p(v), B, S, heads, heads_kv,
              _stream())
    return q, k, v


def pack_swiglu_weight(gate_w, up_w):
    """[gate; up] rows interleaved in blocks of 128 (the layout x2i_gemm_swiglu reads): [2F, K] with F % 128 == 0."""
    F, K = gate_w.shape
    if up_w.shape != gate_w.shape or F % 128:
        raise _lib.X2IError("pack_swiglu_weight: gate / up must be [F, K] with F % 128 == 0")
    return torch.stack([gate_w.view(F // 128, 128, K), up_w.view(F // 128, 128, K)], dim=1).reshape(2 * F, K).contiguous()


def linear_swiglu(x, w_packed, out=None, act=0):
    """act(x @ Wg.T) * (x @ Wu.T) with w_packed = pack_swiglu_weight(Wg, Wu); act 0 SiLU (SwiGLU), 1 tanh-GELU (GEGLU).  x [..., K] -> [..., F]."""
    _chk(x, "x"); _chk(w_packed, "w_packed")
    M, lda = _rows(x)
    N, K = w_packed.shape
    if out is None:
        out = torch.empty(*x.shape[:-1], N // 2, device=x.device, dtype=BF16)
    _, ldc = _rows(out)
    _lib.call("x2i_gemm_swiglu", _p(x), lda, _p(w_packed), w_packed.stride(0), 0, _p(out), ldc, M, N, K, int(act), _stream())
    return out


def linear_residual(x, weight, residual, out, bias=None):
    """out = residual + x @ weight.T (+ bias): the residual connections of a decoder layer (x2i_gemm_gate_residual with gate = NULL).
    x [M, K]; residual / out: [M, N] row-strided views (e.g. one batch element of a layer slot of the [B, C, S, H] capture buffer)."""
    _chk(x, "x"); _chk(weight, "weight"); _chk(residual, "residual"); _chk(out, "out"); _chk(bias, "bias")
    M, lda = _rows(x)
    N, K = weight.shape
    _, ldr = _rows(residual)
    _, ldc = _rows(out)
    _lib.call("x2i_gemm_gate_residual", _p(x), lda, _p(weight), weight.stride(0), _p(bias), 0, 0, M, _p(residual), ldr, _p(out), ldc, 0, 0,
              M, N, K, _stream())
    return out


def causal_attention(q, k, v, kv_start=None, out=None):
    """Causal grouped-query attention of a left-padded prompt: q [B,H,L,128], k, v [B,Hkv,L,128] -> [B, L, H*128]; kv_start int32 [B]
    = first valid key per batch element (None: 0).  Padded query rows (no visible key) output 0."""
    _chk(q, "q"); _chk(k, "k"); _chk(v, "v"); _chk(kv_start, "kv_start", torch.int32)
    B, H, L, d = q.shape
    Hkv = k.shape[1]
    if d != 128 or not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()) or k.shape != (B, Hkv, L, 128) or v.shape != k.shape:
        raise _lib.X2IError("causal_attention: q [B,H,L,128], k,v [B,Hkv,L,128] contiguous")
    if out is None:
        out = torch.empty(B, L, H * 128, device=q.device, dtype=BF16)
    _chk(out, "out")
    _lib.call("x2i_causal_attention", _p(q), _p(k), _p(v), _p(kv_start), _p(out), out.stride(-2), B, H, Hkv, L, _stream())
    return out
