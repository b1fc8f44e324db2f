"""Alteration scoring on the device: the step after training that turns the logit matrices of
the control and the perturbed condition into the ranked mislocalisation list.

Mirrors ``code/main.py``: ``scaling`` (:15-29), ``mat_merge`` (:32-48, here on in-memory matrices instead of the
``log/*.npy`` files) and the scoring part of ``misloc_protein_record`` (:80-84, :143-175).  Arithmetic runs in the
plagnn kernels (``csrc/scoring.cu``); torch only holds the buffers.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check
from .ops import _p, _stream, on_tensor_device, workspace


def _as_matrix(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.PlagnnError("plagnn kernels run on CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.float64):
        raise _lib.PlagnnError(f"expected float32 or float64, got {x.dtype}")
    if x.dim() != 2:
        raise _lib.PlagnnError("expected a 2-D matrix")
    return x if x.stride(1) == 1 else x.contiguous()


@on_tensor_device
def scaling(logit_mat: torch.Tensor) -> torch.Tensor:
    """code/main.py:15-29 — column min-max, then rows divided by their sum, in the input's precision."""
    x = _as_matrix(logit_mat.detach())
    lib = _lib.load()
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=x.dtype, device=x.device)
    nb = lib.plagnn_scoring_workspace_bytes(rows, cols)
    ws = workspace(nb, x.device, "scoring")
    check(lib.plagnn_scaling(_p(x), int(x.dtype == torch.float64), x.stride(0), rows, cols, _p(out), out.stride(0), None, 0,
                             _p(ws), nb, _stream()), "scaling")
    return out


@on_tensor_device
def mat_merge(mats) -> torch.Tensor:
    """code/main.py:32-48 — mean over the runs of scaling(mat); float64 accumulator like ``np.zeros((N, 12))``."""
    mats = [_as_matrix(m.detach()) for m in mats]
    lib = _lib.load()
    rows, cols = mats[0].shape
    acc = torch.zeros((rows, cols), dtype=torch.float64, device=mats[0].device)
    nb = lib.plagnn_scoring_workspace_bytes(rows, cols)
    ws = workspace(nb, acc.device, "scoring")
    for m in mats:
        assert m.shape == (rows, cols)
        check(lib.plagnn_scaling(_p(m), int(m.dtype == torch.float64), m.stride(0), rows, cols, None, 0, _p(acc), acc.stride(0),
                                 _p(ws), nb, _stream()), "scaling")
    # the reference divides by the literal 100 (it always has 100 runs); here: the number of matrices given
    check(lib.plagnn_divide_f64(_p(acc), acc.stride(0), rows, cols, float(len(mats)), _stream()), "divide_f64")
    return acc


@on_tensor_device
def alteration_rank(normal_mat: torch.Tensor, inter_mat: torch.Tensor):
    """code/main.py:80-84 — returns (normal, inter, diff, order): the two scaled matrices, the relative change
    ``(inter - normal) / normal`` and the flat indices of its entries from the largest to the smallest score
    (NaN first, ties by descending index)."""
    normal = scaling(normal_mat.double() if normal_mat.dtype != torch.float64 else normal_mat)
    inter = scaling(inter_mat.double() if inter_mat.dtype != torch.float64 else inter_mat)
    lib = _lib.load()
    rows, cols = normal.shape
    diff = torch.empty_like(normal)
    order = torch.empty(rows * cols, dtype=torch.int64, device=normal.device)
    nb = lib.plagnn_scoring_workspace_bytes(rows, cols)
    ws = workspace(nb, normal.device, "scoring")
    check(lib.plagnn_alteration_rank(_p(normal), normal.stride(0), _p(inter), inter.stride(0), rows, cols, _p(diff),
                                     diff.stride(0), _p(order), _p(ws), nb, _stream()), "alteration_rank")
    return normal, inter, diff, order


def misloc_records(normal_mat: torch.Tensor, inter_mat: torch.Tensor, top: int | None = None):
    """The "all data" table of ``misloc_protein_record`` (code/main.py:143-175) as arrays: for every ranked entry with
    ``diff != -1`` and a score that is > 0 (localisation gained) or < 0 (lost), in rank order:
    row (protein index), col (compartment index), score, normal score, perturbation score, rank (1-based)."""
    normal, inter, diff, order = alteration_rank(normal_mat, inter_mat)
    cols = diff.shape[1]
    flat = diff.reshape(-1)[order]
    keep = (flat != -1.0) & ((flat > 0) | (flat < 0))           # NaN fails both comparisons, as in the reference
    idx = order[keep]
    if top is not None:
        idx = idx[:top]
    out = {"row": idx // cols, "col": idx % cols, "score": diff.reshape(-1)[idx], "normal": normal.reshape(-1)[idx],
           "perturbation": inter.reshape(-1)[idx], "rank": torch.arange(1, idx.numel() + 1, device=idx.device)}
    return out
