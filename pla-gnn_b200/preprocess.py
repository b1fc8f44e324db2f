"""Preprocessing on the device: the two steps of the offline stage that are loops over edges or
dense N×N passes in the reference.

Mirrors ``code/data_preprocess.py``: ``edge_clustering_coefficients(ppi_net, epsilon=0)`` (:175-214) and
``modify_network_topology(ppi_net, pcc_nor, pcc_inter, thr)`` (:217-257) — same arguments (scipy matrices in, scipy
``coo_matrix`` out, same entry order and values); the ``*_device`` functions are the same steps on CUDA tensors.
Arithmetic runs in the plagnn kernels (``csrc/preprocess.cu``); torch holds the buffers and moves data.  No CPU fallback.
The PPI matrix must be a simple 0/1 matrix (no duplicate entries), which is what ``construct_uniprot_ppi`` builds
(code/data_preprocess.py:83-108); a duplicate raises instead of being summed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check
from .graph import build_csr
from .ops import _p, _stream

_STATUS = {1: "a row of the PPI matrix holds a duplicate entry (or the CSR is not sorted)", 2: "node id out of range",
           4: "output capacity too small"}


def _raise_on(status: torch.Tensor, what: str) -> None:
    s = int(status.item())
    if s:
        raise _lib.PlagnnError(f"{what}: " + "; ".join(msg for bit, msg in _STATUS.items() if s & bit))


def _coo_ids(ppi_net, device):
    m = ppi_net.tocoo()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.PlagnnError("plagnn kernels run on CUDA devices only (no CPU fallback)")
    row = torch.from_numpy(np.ascontiguousarray(m.row, dtype=np.int32)).to(dev)
    col = torch.from_numpy(np.ascontiguousarray(m.col, dtype=np.int32)).to(dev)
    return row, col, int(m.shape[0])


def sorted_csr(row: torch.Tensor, col: torch.Tensor, num_nodes: int):
    """CSR of the COO entries with ascending columns inside each row (``ppi_net.tocsr()``, code/data_preprocess.py:184):
    two stable passes of the device sort (K1), by column, then by row."""
    by_col = build_csr(col, row, num_nodes, add_self_loop=False)
    col_sorted = col.to(torch.int32).index_select(0, by_col.eids.long())
    return build_csr(by_col.indices, col_sorted, num_nodes, add_self_loop=False)


def ecc_device(row: torch.Tensor, col: torch.Tensor, num_nodes: int, epsilon: float = 0.0):
    """Edge clustering coefficients of the COO entries (row, col): returns (ecc_row, ecc_col, ecc_data) on the device in
    the order of the reference's lists (code/data_preprocess.py:205-210)."""
    if not row.is_cuda:
        raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    dev = row.device
    nnz = int(row.numel())
    if nnz == 0:
        z = torch.empty(0, dtype=torch.int32, device=dev)
        return z, z.clone(), torch.empty(0, dtype=torch.float64, device=dev)
    csr = sorted_csr(row, col, num_nodes)
    cap = 2 * nnz
    ecc_row = torch.empty(cap, dtype=torch.int32, device=dev)
    ecc_col = torch.empty(cap, dtype=torch.int32, device=dev)
    ecc_data = torch.empty(cap, dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nb = lib.plagnn_ecc_workspace_bytes(num_nodes)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.plagnn_ecc(_p(csr.indptr), _p(csr.indices), num_nodes, nnz, float(epsilon), _p(ecc_row), _p(ecc_col),
                             _p(ecc_data), cap, _p(n_out), _p(status), _p(ws), nb, _stream()), "ecc")
    _raise_on(status, "edge_clustering_coefficients")
    k = int(n_out.item())
    return ecc_row[:k], ecc_col[:k], ecc_data[:k]


def edge_clustering_coefficients(ppi_net, epsilon=0, device="cuda"):
    """code/data_preprocess.py:175-214 — scipy matrix in, ``coo_matrix`` out (float64 data)."""
    from scipy.sparse import coo_matrix
    row, col, n = _coo_ids(ppi_net, device)
    r, c, d = ecc_device(row, col, n, float(epsilon))
    return coo_matrix((d.cpu().numpy(), (r.cpu().numpy(), c.cpu().numpy())), shape=ppi_net.shape)


def pearson_matrix(expr_gcn, device="cuda") -> torch.Tensor:
    """The numeric tail of ``construct_gcn_matrix`` (code/data_preprocess.py:165-170): ``np.corrcoef(expr_gcn)`` with the
    diagonal and the NaN entries (proteins without expression: constant rows) set to 0, as a dense float64 CUDA matrix —
    the form ``modify_network_topology`` consumes (the reference wraps it in a ``coo_matrix`` and densifies it again there)."""
    x = expr_gcn if isinstance(expr_gcn, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(expr_gcn, dtype=np.float64)))
    x = x.to(device=device, dtype=torch.float64)
    if x.dim() != 2 or x.shape[1] < 2:
        raise _lib.PlagnnError("expected a [proteins x samples] matrix with at least two samples")
    if not x.is_cuda:
        raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
    x = x if x.stride(1) == 1 else x.contiguous()
    lib = _lib.load()
    n, s = x.shape
    out = torch.empty((n, n), dtype=torch.float64, device=x.device)
    nb = lib.plagnn_pearson_workspace_bytes(n, s)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.plagnn_pearson(_p(x), x.stride(0), n, s, _p(out), out.stride(0), _p(ws), nb, _stream()), "pearson")
    return out


def _dense_f64(m, device) -> torch.Tensor:
    if isinstance(m, torch.Tensor):
        t = m
    else:
        t = torch.from_numpy(np.ascontiguousarray(m.toarray() if hasattr(m, "toarray") else np.asarray(m), dtype=np.float64))
    t = t.to(device=device, dtype=torch.float64)
    if t.dim() != 2 or t.shape[0] != t.shape[1]:
        raise _lib.PlagnnError("expected a square matrix")
    return t if t.stride(1) == 1 else t.contiguous()


def diff_moments(pcc_nor: torch.Tensor, pcc_inter: torch.Tensor):
    """(mean, std) of ``pcc_inter - pcc_nor`` (code/data_preprocess.py:236,243-244) as Python floats."""
    if not (pcc_nor.is_cuda and pcc_inter.is_cuda):
        raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
    if pcc_nor.dtype != torch.float64 or pcc_inter.dtype != torch.float64 or pcc_nor.shape != pcc_inter.shape or \
            pcc_nor.dim() != 2 or pcc_nor.stride(1) != 1 or pcc_inter.stride(1) != 1:
        raise _lib.PlagnnError("expected two float64 matrices of the same shape with unit column stride")
    lib = _lib.load()
    rows, cols = pcc_nor.shape
    out = torch.empty(2, dtype=torch.float64, device=pcc_nor.device)
    nb = lib.plagnn_diff_moments_workspace_bytes()
    ws = torch.empty(nb, dtype=torch.uint8, device=pcc_nor.device)
    with torch.cuda.device(pcc_nor.device):
        check(lib.plagnn_diff_moments(_p(pcc_nor), pcc_nor.stride(0), _p(pcc_inter), pcc_inter.stride(0), rows, cols, _p(out),
                                      _p(ws), nb, _stream()), "diff_moments")
    m, s = out.cpu().tolist()
    return m, s


def rewire_device(row: torch.Tensor, col: torch.Tensor, num_nodes: int, pcc_nor: torch.Tensor, pcc_inter: torch.Tensor,
                  l_threshold: float, r_threshold: float):
    """Applies the two thresholds (code/data_preprocess.py:250-253) to the adjacency given as COO entries; returns the new
    adjacency as (row, col) in row-major order (code/data_preprocess.py:255)."""
    if not (row.is_cuda and pcc_nor.is_cuda and pcc_inter.is_cuda):
        raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    dev = row.device
    n = int(num_nodes)
    wpr = (n + 31) // 32
    mask = torch.empty(n * wpr, dtype=torch.int32, device=dev)
    new_mask = torch.empty(n * wpr, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    nb = lib.plagnn_rewire_workspace_bytes(n)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    row = row.to(torch.int32).contiguous()
    col = col.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        check(lib.plagnn_adj_bitmask(_p(row), _p(col), int(row.numel()), n, _p(mask), wpr, _p(status), _stream()), "adj_bitmask")
        check(lib.plagnn_rewire(_p(pcc_nor), pcc_nor.stride(0), _p(pcc_inter), pcc_inter.stride(0), n, _p(mask), _p(new_mask), wpr,
                                float(l_threshold), float(r_threshold), _p(rowptr), _p(n_out), _p(ws), nb, _stream()), "rewire")
        _raise_on(status, "modify_network_topology")
        k = int(n_out.item())
        out_row = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
        out_col = torch.empty(max(k, 1), dtype=torch.int32, device=dev)
        check(lib.plagnn_bitmask_to_coo(_p(new_mask), wpr, n, _p(rowptr), _p(out_row), _p(out_col), _stream()), "bitmask_to_coo")
    return out_row[:k], out_col[:k]


def modify_network_topology(ppi_net, pcc_nor, pcc_inter, thr, device="cuda", thresholds=None):
    """code/data_preprocess.py:217-257 — scipy matrices (or dense arrays / CUDA tensors for the two PCC matrices) in,
    ``coo_matrix`` with int64 ones out.  ``thresholds=(left, right)`` overrides the mean ∓ thr·std of the difference."""
    from scipy.sparse import coo_matrix
    row, col, n = _coo_ids(ppi_net, device)
    nor, inter = _dense_f64(pcc_nor, row.device), _dense_f64(pcc_inter, row.device)
    if nor.shape[0] != n or inter.shape[0] != n:
        raise _lib.PlagnnError("PCC matrices and PPI matrix differ in size")
    if thresholds is None:
        diff_mean, diff_std = diff_moments(nor, inter)
        thresholds = (diff_mean - thr * diff_std, diff_mean + thr * diff_std)    # code/data_preprocess.py:245-246
    r, c = rewire_device(row, col, n, nor, inter, thresholds[0], thresholds[1])
    r, c = r.cpu().numpy(), c.cpu().numpy()
    return coo_matrix((np.ones(r.size, dtype=np.int64), (r, c)), shape=ppi_net.shape)
