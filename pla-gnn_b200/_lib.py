"""ctypes binding of libplagnn.so (the C ABI declared in include/plagnn.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# PLAGNN_LIB_PATH selects another build of the same library (e.g. the diagnostics build, csrc/build.py --diag)
LIB_PATH = os.environ.get("PLAGNN_LIB_PATH") or os.path.join(_PKG, "libplagnn.so")

OK = 0
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_TMA, GEMM_NARROW = 0, 1, 2, 3, 4
REDUCE_SUM, REDUCE_MAX = 0, 1
GEMM_MAX_PAIRS = 2


class PlagnnError(RuntimeError):
    pass


class GemmPair(Structure):
    _fields_ = [("a", c_void_p), ("lda", c_int64), ("a_trans", c_int32),
                ("b", c_void_p), ("ldb", c_int64), ("b_trans", c_int32),
                ("k", c_int64)]


class Gnn32Shape(Structure):
    _fields_ = [("num_nodes", c_int64), ("in_feats", c_int32), ("h1", c_int32), ("h2", c_int32), ("h3", c_int32),
                ("h4", c_int32), ("classes", c_int32), ("indptr", c_void_p), ("indices", c_void_p), ("plan", c_void_p),
                ("plan_counts", c_int64 * 3)]


class AdamTensor(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64)]


# name -> (restype, argtypes); every symbol declared in include/plagnn.h appears here
PROTOTYPES = {
    "plagnn_version": (c_int, []),
    "plagnn_launch_count": (ctypes.c_longlong, []),
    "plagnn_profile_enable": (c_int, [c_int]),
    "plagnn_profile_report": (c_size_t, [c_char_p, c_size_t]),
    "plagnn_last_error": (c_char_p, []),
    "plagnn_device_supported": (c_int, []),
    "plagnn_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "plagnn_csr_build": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_plan_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "plagnn_spmm_plan_build": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_size_t, POINTER(c_int64), c_void_p]),
    "plagnn_spmm_partial_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "plagnn_spmm_max_fwd": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_void_p, c_int64, c_int64,
                                    c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_max_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                    c_void_p, c_int64, c_int64, c_void_p]),
    "plagnn_spmm_max_bwd_gather": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_void_p, c_int64,
                                           c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                           c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_sum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_void_p, c_void_p,
                                c_void_p, c_int64, c_int64, c_void_p, c_int, c_float, c_float, c_uint64,
                                c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_sum_slab": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_int64, c_void_p, c_int, c_float, c_void_p, c_int64, c_int,
                                     c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_max_slab": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_void_p, c_int64, c_int64,
                                     c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_plan_range": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(c_int64), c_void_p]),
    "plagnn_spmm_sum_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), c_int64,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_float,
                                     c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_spmm_max_fwd_rows": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), c_int64, c_void_p, c_int64,
                                         c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_dropout_scale": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_float, c_uint64, c_void_p]),
    "plagnn_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "plagnn_gemm": (c_int, [c_int64, c_int64, c_int32, POINTER(GemmPair), c_void_p, c_int, c_float,
                            c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_size_t, c_int, c_void_p]),
    "plagnn_gemm_wgrad_bias_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "plagnn_gemm_wgrad_bias": (c_int, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "plagnn_colsum_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "plagnn_colsum": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "plagnn_act_backward": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_float,
                                    c_void_p, c_void_p, c_int64, c_void_p]),
    "plagnn_bce_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "plagnn_bce_weighted": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                    c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_adam_multi": (c_int, [c_void_p, c_int32, c_int64, c_double, c_double, c_double, c_double, c_double, c_double,
                                  c_void_p]),
    "plagnn_scale_by_device_scalar": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "plagnn_adam_multi_devstep": (c_int, [c_void_p, c_int32, c_int64, c_double, c_double, c_double, c_double, c_void_p, c_void_p,
                                          c_void_p]),
    "plagnn_loc_correction_workspace_bytes": (c_size_t, [c_int64]),
    "plagnn_loc_correction": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64, c_void_p, c_size_t,
                                      c_void_p]),
    "plagnn_scoring_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "plagnn_scaling": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                               c_size_t, c_void_p]),
    "plagnn_divide_f64": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_double, c_void_p]),
    "plagnn_alteration_rank": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "plagnn_pearson_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "plagnn_pearson": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "plagnn_ecc_workspace_bytes": (c_size_t, [c_int64]),
    "plagnn_ecc": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "plagnn_diff_moments_workspace_bytes": (c_size_t, []),
    "plagnn_diff_moments": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    "plagnn_adj_bitmask": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "plagnn_rewire_workspace_bytes": (c_size_t, [c_int64]),
    "plagnn_rewire": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_double, c_double,
                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "plagnn_bitmask_to_coo": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "plagnn_gnn32_arena_bytes": (c_size_t, [POINTER(Gnn32Shape)]),
    "plagnn_gnn32_forward": (c_int, [POINTER(Gnn32Shape), c_void_p, c_int64, POINTER(c_void_p), c_void_p, c_size_t,
                                     c_void_p, c_int64, c_void_p]),
    "plagnn_gnn32_backward": (c_int, [POINTER(Gnn32Shape), c_void_p, c_int64, POINTER(c_void_p), c_void_p, c_size_t,
                                      c_void_p, c_int64, c_void_p, c_int64, POINTER(c_void_p), c_void_p, c_int64,
                                      c_void_p]),
    "plagnn_nccl_available": (c_int, []),
    "plagnn_nccl_get_unique_id": (c_int, [c_void_p]),
    "plagnn_nccl_comm_init": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
    "plagnn_nccl_comm_destroy": (c_int, [c_void_p]),
    "plagnn_nccl_allgather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "plagnn_nccl_reducescatter_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "plagnn_nccl_allreduce": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "plagnn_nccl_alltoall_blocks": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "plagnn_cols_pack": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "plagnn_cols_unpack": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "plagnn_p2p_create": (c_int, [c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "plagnn_p2p_attach": (c_int, [c_void_p, c_void_p]),
    "plagnn_p2p_window": (c_void_p, [c_void_p]),
    "plagnn_p2p_error": (ctypes.c_longlong, [c_void_p]),
    "plagnn_p2p_destroy": (c_int, [c_void_p]),
    "plagnn_p2p_send": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_size_t, ctypes.c_longlong, c_void_p]),
    "plagnn_p2p_send_part": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_int, c_size_t,
                                     ctypes.c_longlong, c_void_p]),
    "plagnn_p2p_wait": (c_int, [c_void_p, ctypes.c_longlong, c_void_p]),
    "plagnn_pad_copy": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "plagnn_transpose": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libplagnn.so and bind every prototype.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlagnnError(
            f"{LIB_PATH} not found: build it with `python pla-gnn_b200/csrc/build.py` "
            "(there is no CPU / PyTorch fallback for the plagnn kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().plagnn_last_error()
        raise PlagnnError(f"{what or 'plagnn call'} failed (code {rc}): {msg.decode() if msg else '?'}")
