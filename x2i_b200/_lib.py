"""ctypes binding of libx2i_b200.so (the C ABI declared in include/x2i_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libx2i_b200.so")

_vp, _i64, _i, _f, _d = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double

class GemmDesc(ctypes.Structure):
    """struct x2i_gemm_desc (include/x2i_b200.h), field for field."""
    _fields_ = [("kind", _i), ("act", _i), ("M", _i), ("N", _i), ("K", _i), ("rows_per_batch", _i), ("heads", _i),
                ("row_offset", _i), ("L_total", _i), ("eps", _f),
                ("A", _vp), ("W", _vp), ("bias", _vp), ("lda", _i64), ("ldw", _i64), ("C", _vp), ("ldc", _i64),
                ("gate", _vp), ("residual", _vp), ("gate_stride", _i64), ("ldr", _i64), ("aux", _vp), ("ldaux", _i64),
                ("rms_q", _vp), ("rms_k", _vp), ("rope", _vp), ("q", _vp), ("k", _vp), ("v", _vp), ("mlp", _vp),
                ("ldmlp", _i64), ("qk_pre", _vp), ("ldqk", _i64), ("mlp_pre", _vp), ("ldmlp_pre", _i64), ("aux_act", _i)]


GEMM_BIAS_ACT, GEMM_GATE_RESIDUAL, GEMM_QKV_ROPE = 0, 1, 2

# name -> argtypes; mirrors include/x2i_b200.h one to one (tests/test_abi.py checks both directions)
SIGNATURES = {
    "x2i_gemm_bias_act": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_gemm_bias_dual": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_gemm_gate_residual": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_gemm_qkv_rope": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "x2i_gemm_grouped": [ctypes.POINTER(GemmDesc), _i, _vp],
    "x2i_gemm_kn": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _vp],
    "x2i_mmdit_attention": [_vp, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _vp],
    "x2i_cross_attention": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_layernorm_affine": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _f, _vp],
    "x2i_add_pos2d": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "x2i_ln_modulate": [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp],
    "x2i_ln_modulate2": [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp],
    "x2i_gate_residual": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_skinny_linear": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "x2i_timestep_sinusoid": [_vp, _vp, _i, _i, _vp],
    "x2i_rope_table": [_vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp],
    "x2i_euler_step": [_vp, _vp, _f, _i64, _vp],
    "x2i_kd_loss_fwd": [_vp, _vp, _i64, _i, _f, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "x2i_kd_loss_bwd": [_vp, _vp, _i64, _i, _f, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "x2i_proj_mix_ln": [_vp, _i, _vp, _f, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp],
    "x2i_mean_over_s": [_vp, _vp, _i, _i, _i, _vp],
    # ---- backward / training
    "x2i_gemm_dgrad": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_gemm_wgrad": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_gemm_bias_act_save": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_gemm_qkv_rope_save": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64,
                               _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "x2i_mmdit_attention_lse": [_vp, _vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _i, _i, _i, _vp],
    "x2i_mmdit_attention_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "x2i_attention_bwd_prep": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp],
    "x2i_qk_norm_rope_bwd": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _f, _vp],
    "x2i_gate_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_ln_modulate_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i, _i, _i, _f, _i, _vp],
    "x2i_colsum": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _vp],
    "x2i_skinny_linear_t": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "x2i_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "x2i_proj_mix_ln_tc": [_vp, _vp, _f, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "x2i_proj_mix_ln_save": [_vp, _i, _vp, _f, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _vp],
    "x2i_mean_over_s_bwd": [_vp, _vp, _i, _i, _i, _vp],
    "x2i_proj_mix_wgrad": [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp],
    # ---- ControlNeXt
    "x2i_conv2d_nhwc": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "x2i_conv_first": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "x2i_groupnorm_nhwc": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp],
    "x2i_conv2d_nhwc_grouped": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "x2i_conv_first_grouped": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "x2i_groupnorm_nhwc_grouped": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp],
    "x2i_gemm_wgrad_splitk": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _i64, _vp],
    "x2i_relu_bwd": [_vp, _vp, _vp, _i64, _vp],
    "x2i_silu": [_vp, _vp, _i64, _vp],
    "x2i_im2col_nhwc": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "x2i_conv2d_nhwc_wgrad": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "x2i_groupnorm_nhwc_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp],
    # ---- VAE decoder
    "x2i_gemm_f32": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp],
    "x2i_softmax_rows": [_vp, _i64, _vp, _i64, _i, _i, _vp],
    "x2i_upsample2x_nhwc": [_vp, _vp, _i, _i, _i, _i, _vp],
    # ---- MLLM prefill
    "x2i_gather_rows": [_vp, _vp, _i64, _i, _vp, _i64, _i64, _i, _i, _i, _vp],
    "x2i_rmsnorm": [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _i, _i, _i, _f, _vp],
    "x2i_rope_half_split": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "x2i_gemm_swiglu": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_softmax_rows_bias": [_vp, _i64, _vp, _i64, _i, _vp, _i64, _i, _i, _vp],
    "x2i_causal_attention": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp],
}
# helpers that return a size instead of a status code
SIZE_FUNCS = {
    "x2i_colsum_workspace_floats": [_i, _i, _i],
    "x2i_skinny_linear_t_workspace_floats": [_i, _i],
    "x2i_proj_mix_wgrad_workspace_floats": [_i, _i, _i],
    "x2i_groupnorm_workspace_floats": [_i, _i, _i],
    "x2i_groupnorm_bwd_workspace_floats": [_i, _i, _i, _i],
    "x2i_gemm_wgrad_workspace_floats": [_i, _i, _i],
    "x2i_proj_mix_ln_tc_supported": [_i, _i, _i, _i],
    "x2i_proj_mix_ln_tc_workspace_floats": [_i, _i, _i, _i],
    "x2i_conv2d_nhwc_wgrad_supported": [_i, _i, _i, _i, _i, _i, _i, _i, _i],
    "x2i_conv2d_nhwc_wgrad_workspace_floats": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
}



_lib = None


class X2IError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise X2IError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(x2i_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        for name, argtypes in SIZE_FUNCS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int64
        L.x2i_version.restype = ctypes.c_int
        L.x2i_last_error.restype = ctypes.c_char_p
        L.x2i_launch_count.restype = ctypes.c_longlong
        _lib = L
    return _lib


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise X2IError(f"{name} failed ({rc}): {lib().x2i_last_error().decode()}")


_graph_replayed = 0  # kernels executed through CUDA-graph replays (each replay re-runs the captured launches)


def note_graph_replay(n_kernels: int) -> None:
    global _graph_replayed
    _graph_replayed += n_kernels


def launch_count() -> int:
    """Kernels of libx2i_b200.so executed so far: direct launches + launches re-run by graph replays."""
    return int(lib().x2i_launch_count()) + _graph_replayed
