"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/plagnn.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "plagnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plagnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import plagnn_b200
    from plagnn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/plagnn.h but not exported by libplagnn.so"


def test_ctypes_prototypes_cover_the_header():
    from plagnn_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_symbols()
    lib = _lib.load()
    assert lib.plagnn_version() >= 100
    assert lib.plagnn_last_error() is not None


def test_workspace_size_queries_are_pure_host_calls():
    from plagnn_b200 import _lib
    lib = _lib.load()
    assert lib.plagnn_csr_build_workspace_bytes(24041, 1400000, 1) > 4 * 4 * (1400000 + 24041)
    assert lib.plagnn_spmm_plan_bytes(24041, 1424041, 128) > 4 * 24041
    assert lib.plagnn_spmm_partial_bytes(0, 503, 1) == 0
    assert lib.plagnn_spmm_partial_bytes(10, 503, 1) >= 2 * 10 * 504 * 4
    assert lib.plagnn_gemm_workspace_bytes(400, 503, 24041) > 0
    assert lib.plagnn_gemm_workspace_bytes(24041, 503, 503) == 0
    assert lib.plagnn_bce_workspace_bytes(8000, 12) >= 32 * 12 * 8


def test_python_constants_equal_the_header_enums():
    """Backend / activation / reducer codes of `_lib.py` are the header's (a drifted constant would pick another kernel)."""
    import re
    from plagnn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "plagnn.h")).read()
    enums = dict((k, int(v)) for k, v in re.findall(r"(PLAGNN_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr))
    for name in ("GEMM_AUTO", "GEMM_SIMT", "GEMM_TCGEN05", "GEMM_TMA", "GEMM_NARROW", "ACT_NONE", "ACT_RELU", "ACT_LEAKY",
                 "ACT_SIGMOID", "REDUCE_MAX", "REDUCE_SUM"):
        assert getattr(_lib, name) == enums["PLAGNN_" + name], name


def test_weight_gradient_workspace_covers_the_narrow_kernel():
    from plagnn_b200 import _lib
    lib = _lib.load()
    # m <= 16: one 12 x 101 partial per 128-row block of the contraction
    assert lib.plagnn_gemm_wgrad_bias_workspace_bytes(12, 100, 24041) >= 188 * 12 * 101 * 4
    assert lib.plagnn_gemm_wgrad_bias_workspace_bytes(400, 503, 24041) >= lib.plagnn_gemm_workspace_bytes(400, 504, 24041)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pla-gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_no_cpu_fallback_raises_without_cuda_tensor():
    import torch
    import plagnn_b200 as P
    with pytest.raises(P.PlagnnError):
        P.ops.aligned(torch.zeros(4, 4))


def test_argument_validation_is_host_side():
    """Entry points reject bad arguments before any CUDA call (return code + message, no exception, no GPU needed)."""
    from plagnn_b200 import _lib
    lib = _lib.load()
    ERR_ARG = -1
    assert lib.plagnn_ecc(None, None, 10, 5, 0.0, None, None, None, 10, None, None, None, 0, None) == ERR_ARG
    assert b"ecc" in lib.plagnn_last_error()
    assert lib.plagnn_diff_moments(None, 4, None, 4, 4, 4, None, None, 0, None) == ERR_ARG
    assert b"diff_moments" in lib.plagnn_last_error()
    assert lib.plagnn_adj_bitmask(None, None, 3, 40, None, 1, None, None) == ERR_ARG
    assert lib.plagnn_rewire(None, 8, None, 8, 8, None, None, 1, -1.0, 1.0, None, None, None, 0, None) == ERR_ARG
    assert b"rewire" in lib.plagnn_last_error()
    assert lib.plagnn_bitmask_to_coo(None, 1, 8, None, None, None, None) == ERR_ARG
    assert lib.plagnn_scaling(None, 0, 12, 10, 12, None, 12, None, 12, None, 0, None) == ERR_ARG
    assert lib.plagnn_divide_f64(None, 12, 10, 12, 100.0, None) == ERR_ARG
