"""Thin torch-tensor wrappers over the C ABI (include/plagnn.h).

PyTorch is used here for device memory and streams only; every computation below is a call into
libplagnn.so on the current CUDA stream.  All functions raise on failure (no fallback path).
"""
from __future__ import annotations

import ctypes
import functools
import math

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, GEMM_AUTO, GEMM_NARROW, GEMM_SIMT, GEMM_TCGEN05, GEMM_TMA, REDUCE_MAX,
                   REDUCE_SUM, GemmPair, check)

LEAKY_SLOPE = 0.01          # F.leaky_relu default, code/model.py:21,23,25,27
ROW_ALIGN = 32              # floats: rows of internal matrices start on 128-byte boundaries


class KernelTimer:
    """Optional CUDA-event timing of individual kernel calls on the launching stream (bench.py roofline)."""

    def __init__(self):
        self.records = {}

    def add(self, key, start, end):
        self.records.setdefault(key, []).append((start, end))

    def summary(self):
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v) / len(v)) for k, v in self.records.items()}


TIMER: KernelTimer | None = None


class _timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if TIMER is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if TIMER is not None:
            self.e.record()
            TIMER.add(self.key, self.s, self.e)
        return False


def profile_start(aggregation_only: bool = False) -> None:
    """Start per-call CUDA-event timing inside the library (all entry points, C- or Python-orchestrated).
    aggregation_only: record the spmm_* entry points only (6 calls per epoch: negligible overhead)."""
    _lib.load().plagnn_profile_enable(2 if aggregation_only else 1)


def profile_stop() -> dict:
    """Stop and return {(name, t0, t1, t2): (calls, total_ms)}."""
    lib = _lib.load()
    lib.plagnn_profile_enable(0)
    need = lib.plagnn_profile_report(None, 0)
    buf = ctypes.create_string_buffer(need + 16)
    lib.plagnn_profile_report(buf, need + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        name, t0, t1, t2, calls, ms = line.split()
        out[(name, int(t0), int(t1), int(t2))] = (int(calls), float(ms))
    return out


def launch_count() -> int:
    return int(_lib.load().plagnn_launch_count())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_of(obj, depth: int = 0):
    """CUDA device of the first tensor found in obj (tensor, object with .indptr, or a short list / tuple of those)."""
    if isinstance(obj, torch.Tensor):
        return obj.device if obj.is_cuda else None
    ip = getattr(obj, "indptr", None)
    if isinstance(ip, torch.Tensor):
        return ip.device if ip.is_cuda else None
    if depth < 2 and isinstance(obj, (list, tuple)):
        for o in obj:
            d = _device_of(o, depth + 1)
            if d is not None:
                return d
    return None


def on_tensor_device(fn):
    """Runs the wrapped entry point with the CUDA device of its tensor arguments current, so that the stream handed to the
    library, the kernels' context and the cached TMA tensor maps belong to the GPU that holds the data — the reference's
    `-d cuda:1` (code/main_normal.py:30,66: g.to(device), model.to(device)) while device 0 is current."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in args:
            dev = _device_of(a)
            if dev is not None:
                break
        if dev is None:
            for a in kwargs.values():
                dev = _device_of(a)
                if dev is not None:
                    break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _p(t: torch.Tensor | None) -> ctypes.c_void_p | None:
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def pitch_of(cols: int) -> int:
    return (cols + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


def alloc(rows: int, cols: int, device, dtype=torch.float32, zero: bool = False) -> torch.Tensor:
    """rows x cols matrix whose rows are 128-byte aligned (a [:, :cols] view of a padded buffer)."""
    pitch = pitch_of(cols)
    buf = (torch.zeros if zero else torch.empty)((rows, pitch), device=device, dtype=dtype)
    return buf[:, :cols]


def is_aligned(x: torch.Tensor) -> bool:
    return (x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.stride(0) >= (x.shape[1] + 3) // 4 * 4
            and x.data_ptr() % 16 == 0)


def _require_cuda_f32(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise _lib.PlagnnError(f"expected float32, got {t.dtype}")


@on_tensor_device
def aligned(x: torch.Tensor) -> torch.Tensor:
    """Returns x if its rows are 16-byte aligned, else a padded copy made by plagnn_pad_copy."""
    _require_cuda_f32(x)
    if is_aligned(x):
        return x
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.contiguous()
    out = alloc(x.shape[0], x.shape[1], x.device)
    check(_lib.load().plagnn_pad_copy(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), out.stride(0), _stream()),
          "pad_copy")
    return out


@on_tensor_device
def transpose(x: torch.Tensor) -> torch.Tensor:
    """x^T as a new row-aligned matrix (used for the K-contiguous weight operand of the input-gradient GEMMs)."""
    _require_cuda_f32(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = alloc(x.shape[1], x.shape[0], x.device)
    check(_lib.load().plagnn_transpose(_p(x), x.shape[0], x.shape[1], x.stride(0), _p(out), out.stride(0), _stream()),
          "transpose")
    return out


# ------------------------------------------------------------------------------------------------
# workspace cache (per device + stream; reuse is ordered by the stream)
# ------------------------------------------------------------------------------------------------
_WS: dict = {}


def workspace(nbytes: int, device, tag: str = "ws") -> torch.Tensor | None:
    if nbytes <= 0:
        return None
    key = (str(device), torch.cuda.current_stream().cuda_stream, tag)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------
# K3 dense contraction
# ------------------------------------------------------------------------------------------------
@on_tensor_device
def gemm(m: int, n: int, pairs, bias=None, act=ACT_NONE, gate=None, gate_act=ACT_NONE, out=None,
         slope: float = LEAKY_SLOPE, backend: int = GEMM_AUTO) -> torch.Tensor:
    """C[m x n] = epilogue(sum_p op(A_p) op(B_p)).  pairs: list of (a, a_trans, b, b_trans, k)."""
    lib = _lib.load()
    dev = pairs[0][0].device
    arr = (GemmPair * len(pairs))()
    ktot = 0
    for i, (a, a_trans, b, b_trans, k) in enumerate(pairs):
        _require_cuda_f32(a, b)
        assert a.stride(1) == 1 and b.stride(1) == 1
        arr[i] = GemmPair(a.data_ptr(), a.stride(0), int(a_trans), b.data_ptr(), b.stride(0), int(b_trans), int(k))
        ktot += int(k)
    if out is None:
        out = alloc(m, n, dev)
    _require_cuda_f32(out, bias, gate)
    ws_bytes = lib.plagnn_gemm_workspace_bytes(m, n, ktot)
    ws = workspace(ws_bytes, dev, "gemm")
    with _timed(("gemm", m, n, ktot)):
        check(lib.plagnn_gemm(m, n, len(pairs), arr, _p(bias), act, slope, _p(gate),
                              gate.stride(0) if gate is not None else 0, gate_act, _p(out), out.stride(0), _p(ws),
                              ws_bytes if ws is not None else 0, backend, _stream()), "gemm")
    return out


@on_tensor_device
def gemm_wgrad_bias(dz: torch.Tensor, x: torch.Tensor, dw: torch.Tensor | None = None, db: torch.Tensor | None = None):
    """dW[m x n] = dz^T x and db[m] = dz.sum(0) in one pass (dz: [k x m], x: [k x n], rows 16-byte aligned)."""
    lib = _lib.load()
    _require_cuda_f32(dz, x, dw, db)
    k, m = dz.shape
    n = x.shape[1]
    assert x.shape[0] == k and dz.stride(1) == 1 and x.stride(1) == 1
    if dw is None:
        dw = torch.empty((m, n), device=dz.device, dtype=torch.float32)
    if db is None:
        db = torch.empty(m, device=dz.device, dtype=torch.float32)
    nb = lib.plagnn_gemm_wgrad_bias_workspace_bytes(m, n, k)
    ws = workspace(nb, dz.device, "gemm")
    with _timed(("gemm", m, n, k)):
        check(lib.plagnn_gemm_wgrad_bias(m, n, _p(dz), dz.stride(0), _p(x), x.stride(0), k, _p(dw), dw.stride(0), _p(db),
                                         _p(ws), nb if ws is not None else 0, _stream()), "gemm_wgrad_bias")
    return dw, db


@on_tensor_device
def colsum(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda_f32(x)
    rows, cols = x.shape
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    nb = lib.plagnn_colsum_workspace_bytes(rows, cols)
    ws = workspace(nb, x.device, "colsum")
    check(lib.plagnn_colsum(_p(x), rows, cols, x.stride(0), _p(out), _p(ws), nb, _stream()), "colsum")
    return out


@on_tensor_device
def act_backward(dy: torch.Tensor, y: torch.Tensor | None, act: int, slope: float = LEAKY_SLOPE,
                 row_scale: torch.Tensor | None = None) -> torch.Tensor:
    """dz = dy * act'(y) * row_scale[:, None]   (y / row_scale optional)."""
    _require_cuda_f32(dy, y, row_scale)
    if dy.stride(1) != 1:
        dy = dy.contiguous()
    out = alloc(dy.shape[0], dy.shape[1], dy.device)
    check(_lib.load().plagnn_act_backward(_p(dy), dy.stride(0), _p(y), y.stride(0) if y is not None else 0, dy.shape[0],
                                          dy.shape[1], act, slope, _p(row_scale), _p(out), out.stride(0), _stream()),
          "act_backward")
    return out


def row_scale(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return act_backward(x, None, ACT_NONE, row_scale=scale)


# ------------------------------------------------------------------------------------------------
# K2 aggregation
# ------------------------------------------------------------------------------------------------
@on_tensor_device
def spmm_max_fwd(csc, x: torch.Tensor):
    """csc: graph.Csr (in-edge structure + plan).  Returns (out, arg) with x's row pitch."""
    lib = _lib.load()
    x = aligned(x)
    n, f = csc.num_rows, x.shape[1]
    out = alloc(n, f, x.device)
    arg = alloc(n, f, x.device, dtype=torch.int32)
    nb = lib.plagnn_spmm_partial_bytes(csc.counts[2], f, REDUCE_MAX)
    part = workspace(nb, x.device, "spmm_partial")
    with _timed(("spmm_max_fwd", f)):
        check(lib.plagnn_spmm_max_fwd(_p(csc.indptr), _p(csc.indices), _p(csc.plan), csc.counts_c, n, _p(x), x.stride(0),
                                      f, _p(out), _p(arg), out.stride(0), _p(part), nb, _stream()), "spmm_max_fwd")
    return out, arg


@on_tensor_device
def spmm_max_bwd(dz: torch.Tensor, arg: torch.Tensor, z: torch.Tensor | None, n_src: int) -> torch.Tensor:
    """dx[arg[v,f], f] += dz[v,f] (* (z>0) when z is given)."""
    lib = _lib.load()
    dz = aligned(dz)
    f = dz.shape[1]
    dx = alloc(n_src, f, dz.device)
    with _timed(("spmm_max_bwd", f)):
        check(lib.plagnn_spmm_max_bwd(_p(dz), dz.stride(0), _p(arg), arg.stride(0), _p(z),
                                      z.stride(0) if z is not None else 0, dz.shape[0], f, _p(dx), n_src, dx.stride(0),
                                      _stream()), "spmm_max_bwd")
    return dx


@on_tensor_device
def spmm_max_bwd_gather(csr, dz: torch.Tensor, arg: torch.Tensor, z: torch.Tensor | None) -> torch.Tensor:
    """Ordered twin of spmm_max_bwd over the out-edge structure `csr` (graph must have no duplicate edges)."""
    lib = _lib.load()
    dz = aligned(dz)
    n_src, f = csr.num_rows, dz.shape[1]
    dx = alloc(n_src, f, dz.device)
    nb = lib.plagnn_spmm_partial_bytes(csr.counts[2], f, REDUCE_SUM)
    part = workspace(nb, dz.device, "spmm_partial")
    with _timed(("spmm_max_bwd_gather", f)):
        check(lib.plagnn_spmm_max_bwd_gather(_p(csr.indptr), _p(csr.indices), _p(csr.plan), csr.counts_c, n_src, _p(dz),
                                             dz.stride(0), _p(arg), arg.stride(0), _p(z),
                                             z.stride(0) if z is not None else 0, f, _p(dx), dx.stride(0), _p(part), nb,
                                             _stream()), "spmm_max_bwd_gather")
    return dx


@on_tensor_device
def spmm_sum(csx, x: torch.Tensor, w=None, scale=None, bias=None, act=ACT_NONE, dropout_p: float = 0.0,
             dropout_seed: int = 0, slope: float = LEAKY_SLOPE, w_in_csr_order: bool = False) -> torch.Tensor:
    """w: one weight per edge, indexed by edge id (DGL's convention) or, with w_in_csr_order, by position in csx."""
    lib = _lib.load()
    x = aligned(x)
    n, f = csx.num_rows, x.shape[1]
    out = alloc(n, f, x.device)
    nb = lib.plagnn_spmm_partial_bytes(csx.counts[2], f, REDUCE_SUM)
    part = workspace(nb, x.device, "spmm_partial")
    with _timed(("spmm_sum", f)):
        check(lib.plagnn_spmm_sum(_p(csx.indptr), _p(csx.indices), _p(csx.eids) if (w is not None and not w_in_csr_order) else None,
                                  _p(csx.plan), csx.counts_c, n, _p(w), _p(scale), _p(x), x.stride(0), f, _p(bias), act,
                                  slope, float(dropout_p), int(dropout_seed), _p(out), out.stride(0), _p(part), nb,
                                  _stream()), "spmm_sum")
    return out


@on_tensor_device
def spmm_sum_slabs(slabs, x: torch.Tensor, ws=None, scale=None, bias=None, act=ACT_NONE, slope: float = LEAKY_SLOPE) -> torch.Tensor:
    """The sum reducer as one pass per SOURCE slab (plagnn_spmm_sum_slab): `slabs` = CSR structures of the same rows whose
    entries are restricted to consecutive source ranges, `ws` = edge weights of each in its CSR order (or None)."""
    lib = _lib.load()
    x = aligned(x)
    n, f = slabs[0].num_rows, x.shape[1]
    out = alloc(n, f, x.device)
    nb = max(lib.plagnn_spmm_partial_bytes(c.counts[2], f, REDUCE_SUM) for c in slabs)
    part = workspace(nb, x.device, "spmm_partial")
    for i, csx in enumerate(slabs):
        w = None if ws is None else ws[i]
        last = i + 1 == len(slabs)
        check(lib.plagnn_spmm_sum_slab(_p(csx.indptr), _p(csx.indices), None, _p(csx.plan), csx.counts_c, n, _p(w), _p(scale),
                                       _p(x), x.stride(0), f, _p(bias), act, slope, _p(out) if i else None, out.stride(0),
                                       int(last), _p(out), out.stride(0), _p(part), nb, _stream()), "spmm_sum_slab")
    return out


@on_tensor_device
def spmm_max_slabs(slabs, x: torch.Tensor):
    """The max reducer as one pass per source slab (plagnn_spmm_max_slab).  Returns (out, arg); on equal values the earlier
    slab wins, inside a slab the first in-edge."""
    lib = _lib.load()
    x = aligned(x)
    n, f = slabs[0].num_rows, x.shape[1]
    out = alloc(n, f, x.device)
    arg = alloc(n, f, x.device, dtype=torch.int32)
    nb = max(lib.plagnn_spmm_partial_bytes(c.counts[2], f, REDUCE_MAX) for c in slabs)
    part = workspace(nb, x.device, "spmm_partial")
    for i, csx in enumerate(slabs):
        last = i + 1 == len(slabs)
        check(lib.plagnn_spmm_max_slab(_p(csx.indptr), _p(csx.indices), _p(csx.plan), csx.counts_c, n, _p(x), x.stride(0), f,
                                       _p(out) if i else None, _p(arg) if i else None, out.stride(0), int(last), _p(out), _p(arg),
                                       out.stride(0), _p(part), nb, _stream()), "spmm_max_slab")
    return out, arg


@on_tensor_device
def plan_range(csx, row_begin: int, row_end: int):
    """Host handle (ctypes int64[4]) for aggregating only rows [row_begin, row_end) of `csx` (set-up call)."""
    rng = (ctypes.c_int64 * 4)()
    check(_lib.load().plagnn_spmm_plan_range(_p(csx.plan), csx.num_rows, row_begin, row_end, rng, _stream()), "spmm_plan_range")
    return rng


@on_tensor_device
def spmm_sum_rows(csx, rng, x: torch.Tensor, out: torch.Tensor, w=None, scale=None, bias=None, act=ACT_NONE,
                  slope: float = LEAKY_SLOPE, w_in_csr_order: bool = False) -> torch.Tensor:
    """Row-range aggregation into the rows of a preallocated `out` (other rows untouched)."""
    lib = _lib.load()
    f = x.shape[1]
    nb = lib.plagnn_spmm_partial_bytes(csx.counts[2], f, REDUCE_SUM)
    part = workspace(nb, x.device, "spmm_partial")
    check(lib.plagnn_spmm_sum_rows(_p(csx.indptr), _p(csx.indices), _p(csx.eids) if (w is not None and not w_in_csr_order) else None, _p(csx.plan),
                                   csx.counts_c, rng, csx.num_rows, _p(w), _p(scale), _p(x), x.stride(0), f, _p(bias), act,
                                   slope, _p(out), out.stride(0), _p(part), nb, _stream()), "spmm_sum_rows")
    return out


@on_tensor_device
def spmm_max_fwd_rows(csc, rng, x: torch.Tensor, out: torch.Tensor, arg: torch.Tensor):
    """Row-range max aggregation into the rows of preallocated `out` / `arg` (other rows untouched)."""
    lib = _lib.load()
    f = x.shape[1]
    nb = lib.plagnn_spmm_partial_bytes(csc.counts[2], f, REDUCE_MAX)
    part = workspace(nb, x.device, "spmm_partial")
    check(lib.plagnn_spmm_max_fwd_rows(_p(csc.indptr), _p(csc.indices), _p(csc.plan), csc.counts_c, rng, csc.num_rows, _p(x),
                                       x.stride(0), f, _p(out), _p(arg), out.stride(0), _p(part), nb, _stream()), "spmm_max_fwd_rows")
    return out, arg


@on_tensor_device
def dropout_scale_(grad: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    check(_lib.load().plagnn_dropout_scale(_p(grad), grad.shape[0], grad.shape[1], grad.stride(0), float(p), int(seed),
                                           _stream()), "dropout_scale")
    return grad


# ------------------------------------------------------------------------------------------------
# K4 loss / optimiser / label decision
# ------------------------------------------------------------------------------------------------
@on_tensor_device
def bce_weighted(prob: torch.Tensor, target: torch.Tensor, index: torch.Tensor | None, cw: torch.Tensor,
                 cwp1: torch.Tensor, want_grad: bool = True, grad_scale: float = 1.0):
    """Returns (loss[1], dprob[N x C] or None).  index: int64 device tensor of selected rows or None."""
    lib = _lib.load()
    _require_cuda_f32(prob, target, cw, cwp1)
    assert prob.stride(1) == 1 and target.stride(1) == 1
    n, c = prob.shape
    r = n if index is None else index.numel()
    loss = torch.empty(1, device=prob.device, dtype=torch.float32)
    dprob = alloc(n, c, prob.device) if want_grad else None
    nb = lib.plagnn_bce_workspace_bytes(r, c)
    ws = workspace(nb, prob.device, "bce")
    check(lib.plagnn_bce_weighted(_p(prob), prob.stride(0), _p(target), target.stride(0), _p(index), r, n, c, _p(cw),
                                  _p(cwp1), float(grad_scale), _p(loss), _p(dprob),
                                  dprob.stride(0) if dprob is not None else 0, _p(ws), nb, _stream()), "bce_weighted")
    return loss, dprob


@on_tensor_device
def scale_by_device_scalar(x: torch.Tensor, scalar: torch.Tensor) -> torch.Tensor:
    """x * scalar with the scalar read on the device (0-dim or 1-element fp32 CUDA tensor)."""
    _require_cuda_f32(x, scalar)
    assert x.dim() == 2 and x.stride(1) == 1 and scalar.numel() == 1
    out = alloc(x.shape[0], x.shape[1], x.device)
    check(_lib.load().plagnn_scale_by_device_scalar(_p(x), x.stride(0), x.shape[0], x.shape[1], _p(scalar), _p(out), out.stride(0),
                                                    _stream()), "scale_by_device_scalar")
    return out


@on_tensor_device
def loc_correction(prob: torch.Tensor, alpha: float) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda_f32(prob)
    n, c = prob.shape
    pred = torch.empty((n, c), device=prob.device, dtype=torch.float32)
    nb = lib.plagnn_loc_correction_workspace_bytes(c)
    ws = workspace(nb, prob.device, "loc")
    check(lib.plagnn_loc_correction(_p(prob), prob.stride(0), n, c, float(alpha), _p(pred), pred.stride(0), _p(ws), nb,
                                    _stream()), "loc_correction")
    return pred


@on_tensor_device
def adam_multi(table: torch.Tensor, count: int, max_numel: int, lr: float, beta1: float, beta2: float, eps: float,
               step: int) -> None:
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
    check(_lib.load().plagnn_adam_multi(_p(table), count, max_numel, lr, beta1, beta2, eps, bc1, bc2_sqrt, _stream()),
          "adam_multi")
