"""Fused multi-tensor Adam (``plagnn_adam_multi``): one launch for all parameter tensors.
Replaces ``torch.optim.Adam(model.parameters(), lr)`` + ``.step()`` of ``code/train.py:180,205``
(betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)."""
from __future__ import annotations

import torch

from . import _lib, ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam mirrors the reference's plain Adam (no weight decay / amsgrad)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            rows = []
            max_numel = 0
            step = None
            dev = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise _lib.PlagnnError("FusedAdam needs float32 CUDA parameters (no CPU fallback)")
                if not p.is_contiguous():
                    raise _lib.PlagnnError("FusedAdam needs contiguous parameters")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if step is None:
                    step = st["step"]
                elif step != st["step"]:
                    raise _lib.PlagnnError("FusedAdam: parameters of one group must share the step count")
                st["_grad_keepalive"] = g
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
                max_numel = max(max_numel, p.numel())
                dev = p.device
            if not rows:
                continue
            # The pointer table lives on the device and is re-uploaded only when a pointer changes (the caching
            # allocator hands the gradients the same blocks every epoch).  A pageable .to(device) here would block the
            # host until the whole epoch has drained: one stream sync per step.
            key = tuple(rows)
            if group.get("_table_key") != key:
                host = torch.tensor(rows, dtype=torch.int64).pin_memory()
                group["_table_host"] = (host, group.get("_table_host", (None,))[0])   # keep the previous one alive too
                group["_table"] = host.to(dev, non_blocking=True)
                group["_table_key"] = key
            table = group["_table"]
            with torch.cuda.device(dev):
                ops.adam_multi(table, len(rows), max_numel, group["lr"], group["betas"][0], group["betas"][1],
                               group["eps"], step)
        return loss
