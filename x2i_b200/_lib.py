"""ctypes binding of libx2i_b200.so (the C ABI declared in include/x2i_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libx2i_b200.so")

_vp, _i64, _i, _f, _d = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double

# name -> argtypes; mirrors include/x2i_b200.h one to one (tests/test_abi.py checks both directions)
SIGNATURES = {
    "x2i_gemm_bias_act": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp],
    "x2i_gemm_bias_dual": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_gemm_gate_residual": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_gemm_qkv_rope": [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "x2i_gemm_kn": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _vp],
    "x2i_mmdit_attention": [_vp, _vp, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _vp],
    "x2i_ln_modulate": [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _f, _vp],
    "x2i_gate_residual": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp],
    "x2i_skinny_linear": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "x2i_timestep_sinusoid": [_vp, _vp, _i, _i, _vp],
    "x2i_rope_table": [_vp, _i, _i, _i, _i, _d, _vp, _vp, _vp, _vp],
    "x2i_euler_step": [_vp, _vp, _f, _i64, _vp],
    "x2i_kd_loss_fwd": [_vp, _vp, _i64, _i, _f, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "x2i_kd_loss_bwd": [_vp, _vp, _i64, _i, _f, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp],
    "x2i_proj_mix_ln": [_vp, _i, _vp, _f, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp],
    "x2i_mean_over_s": [_vp, _vp, _i, _i, _i, _vp],
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
