"""On-device label decision and metrics: vectorised replacements for the Python
row loops of ``code/train.py:19-40`` (protein_loc_correction) and ``:43-86`` (performances_record)."""
from __future__ import annotations

import torch

from . import ops


def protein_loc_correction(loc_proba: torch.Tensor, alpha: float) -> torch.Tensor:
    """Column min-max, row-normalise, threshold max-(max-min)*alpha per row -> 0/1 matrix (float64 like the
    reference, same device as the input).  One kernel pair instead of a 24 041-iteration Python loop."""
    p = loc_proba.detach()
    if p.stride(1) != 1:
        p = p.contiguous()
    return ops.loc_correction(p, alpha).double()


def performances_record(loc_true: torch.Tensor, loc_pred: torch.Tensor):
    """AIM / COV / mlACC means over rows (code/train.py:43-86).  Bookkeeping on R x 12 0/1 matrices with
    torch ops on the tensors' device; the three scalars are read back with one transfer."""
    t = loc_true.detach().long() == 1
    p = loc_pred.detach().long() == 1
    inter = (t & p).sum(1).float()
    pred = p.sum(1).float()
    real = t.sum(1).float()
    union = (t | p).sum(1).float()
    aim = torch.where(pred == 0, torch.zeros_like(inter), inter / pred)
    n = len(t)
    out = torch.stack([aim.double().sum(), (inter / real).double().sum(), (inter / union).double().sum()]) / n
    a, c, m = out.tolist()
    return a, c, m
