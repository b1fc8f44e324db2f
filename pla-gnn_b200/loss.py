"""Fused class-weighted clamped BCE of the reference (``code/train.py:89-108``) and the class weights
(``code/train.py:111-126``), as autograd Functions over ``plagnn_bce_weighted``."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def weight_cal(loc_mat) -> np.ndarray:
    """(rows with >= 1 label - class count) / class count, float64[C]  (code/train.py:111-126).
    Host-side, once per run; vectorised instead of the reference's Python row loop."""
    loc_mat = np.asarray(loc_mat)
    class_num = loc_mat.sum(axis=0)
    sample_num = int((loc_mat.sum(axis=1) != 0).sum())
    return (sample_num - class_num) / class_num


_CW_CACHE: dict = {}


def _class_weights(i_weight, device):
    """fp32 device copies of w_i and (w_i + 1): the reference multiplies an fp32 tensor by the float64
    scalar w_i and divides by the float64 scalar (w_i + 1); torch rounds each scalar to fp32 first."""
    w64 = np.asarray(i_weight, dtype=np.float64)
    key = (w64.tobytes(), str(device))
    hit = _CW_CACHE.get(key)
    if hit is None:
        cw = torch.tensor(w64.astype(np.float32), device=device)
        cwp1 = torch.tensor((w64 + 1.0).astype(np.float32), device=device)
        hit = _CW_CACHE[key] = (cw, cwp1)
    return hit


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, target, index, cw, cwp1):
        p = prob.detach()
        if p.stride(1) != 1:
            p = p.contiguous()
        t = target.detach()
        if t.stride(1) != 1:
            t = t.contiguous()
        loss, dprob = ops.bce_weighted(p, t, index, cw, cwp1, want_grad=ctx.needs_input_grad[0])
        ctx.dprob = dprob
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        # d loss / d prob was formed by the loss kernel; times the incoming gradient of the scalar loss (a device scalar)
        return ops.scale_by_device_scalar(ctx.dprob, grad_out), None, None, None, None


def multi_loss(input, target, i_weight):
    """Drop-in for ``train.multi_loss(logits[train_index], labels[train_index], i_weight)``."""
    cw, cwp1 = _class_weights(i_weight, input.device)
    return _BCEFunction.apply(input, target, None, cw, cwp1)


def multi_loss_indexed(logits, labels, index, i_weight):
    """Same loss with the row gather fused in: equals multi_loss(logits[index], labels[index], i_weight);
    the gradient is written as a dense N x C matrix (zeros off the selected rows)."""
    cw, cwp1 = _class_weights(i_weight, logits.device)
    if not isinstance(index, torch.Tensor):
        index = torch.as_tensor(index, dtype=torch.int64)
    index = index.to(device=logits.device, dtype=torch.int64)
    return _BCEFunction.apply(logits, labels, index, cw, cwp1)
