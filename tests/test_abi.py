"""CPU: the C-ABI library builds, loads and exports exactly what include/x2i_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(with_types=False):
    src = open(os.path.join(ROOT, "include", "x2i_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls, rets = {}, {}
    for m in re.finditer(r"\b(int|int64_t|long long|const char\*)\s+(x2i_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        decls[m.group(2)] = 0 if args == "void" else len([a for a in args.split(",") if a.strip()])
        rets[m.group(2)] = m.group(1)
    return (decls, rets) if with_types else decls


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from x2i_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(built):
    decls = _declared()
    assert len(decls) >= 18
    lib = ctypes.CDLL(built.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/x2i_b200.h but not exported"


def test_bindings_match_header(built):
    decls = _declared()
    for name, argtypes in built.SIGNATURES.items():
        assert name in decls, f"{name} bound in _lib.py but not declared in the header"
        assert len(argtypes) == decls[name], f"{name}: {len(argtypes)} bound args vs {decls[name]} declared"
    for name, argtypes in built.SIZE_FUNCS.items():
        assert name in decls and len(argtypes) == decls[name]
    unbound = set(decls) - set(built.SIGNATURES) - set(built.SIZE_FUNCS) - {"x2i_version", "x2i_last_error", "x2i_launch_count"}
    assert not unbound, f"declared but unbound: {unbound}"
    # return types: `int` status codes are bound with restype c_int (SIGNATURES), `int64_t` sizes with c_int64 (SIZE_FUNCS);
    # a status function bound as int64 would read the undefined upper half of RAX
    _, rets = _declared(with_types=True)
    for name in built.SIGNATURES:
        assert rets[name] == "int", f"{name} returns {rets[name]} but is bound as a status (int) function"
    for name in built.SIZE_FUNCS:
        assert rets[name] == "int64_t", f"{name} returns {rets[name]} but is bound as a size (int64) function"


def test_version_and_no_gpu_error_path(built):
    lib = built.lib()
    assert lib.x2i_version() >= 100
    import torch
    if not torch.cuda.is_available():
        # without a device every entry point must fail loudly (no CPU fallback)
        with pytest.raises(built.X2IError):
            built.call("x2i_euler_step", 0, 0, 0.0, 8, 0)


def test_library_has_blackwell_sass(built):
    """tcgen05 / TMA evidence in the shipped SASS (B200_PROFILING.md): UTCHMMA, LDTM, UTMALDG."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built.LIB_PATH], capture_output=True, text=True).stdout
    for op in ("UTCHMMA", "LDTM", "UTMALDG", "STTM"):
        assert op in sass, f"{op} missing from SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path found"


def test_shape_rules_of_the_tensor_pipe_forms(built):
    """The `_supported` helpers are pure host logic (no device needed): which shapes take the implicit convolution weight gradient and the
    tensor-pipe layer-mixing convolution, and which fall back to the explicit / stencil kernels."""
    from x2i_b200 import _lib
    L = _lib.lib()
    wg = L.x2i_conv2d_nhwc_wgrad_supported  # (H, W, Cin, Cout, KH, KW, stride, pad, pad_end)
    assert wg(512, 512, 128, 128, 3, 3, 1, 1, 1) == 1          # ControlNeXt resnet conv at 512 x 512: Wo = 512
    assert wg(512, 512, 128, 128, 3, 3, 2, 1, 1) == 1          # the stride-2 down-sampling conv: Wo = 256
    assert wg(128, 128, 256, 3072, 2, 2, 2, 0, 0) == 1         # the 2x2 / stride-2 output conv: Wo = 64
    assert wg(16, 16, 64, 64, 3, 3, 1, 1, 1) == 1              # short rows: 64 % Wo == 0 and Ho * Wo % 64 == 0 (boxes of 4 rows)
    assert wg(24, 40, 64, 64, 3, 3, 1, 1, 1) == 0              # Wo = 40 does not tile into 64-pixel blocks -> explicit im2col form
    assert wg(12, 16, 64, 64, 3, 3, 1, 1, 1) == 1 and wg(6, 16, 64, 64, 3, 3, 1, 1, 1) == 0   # Ho * Wo must be a multiple of 64
    assert wg(512, 512, 3, 64, 3, 3, 2, 1, 1) == 0             # Cin % 64 (the stem pads its 3 channels before calling)
    assert wg(33, 64, 64, 64, 3, 3, 2, 1, 1) == 0              # stride 2 needs even H and W
    pc = L.x2i_proj_mix_ln_tc_supported  # (B, C, S, H)
    assert pc(1, 37, 512, 2048) == 1 and pc(4, 29, 512, 3584) == 1   # the Qwen2.5-VL 3B / 7B projectors
    assert pc(1, 37, 500, 2048) == 0                           # S % 128 != 0 -> stencil kernel
    assert pc(1, 64, 512, 2048) == 0                           # more than 40 input channels (weights staged in 4 KB of shared memory)
    assert pc(1, 25, 512, 896) == 1 and pc(1, 25, 512, 256) == 0    # InternVL-1B projector width is fine, tiny H is not worth it
