"""The epoch body of the reference's training loop as one replayable unit.

``code/train.py:195-205`` runs, 200 times per model and 100 models per condition, the same fixed-shape sequence::

    optimizer.zero_grad(); logits = model(g, features)
    train_loss = multi_loss(logits[train_index], labels[train_index], i_weight)
    train_loss.backward(); optimizer.step()

``TrainStep`` is that sequence without autograd in the way: four library calls (whole-network forward, loss + loss gradient,
whole-network backward, Adam with its step count on the device) into buffers allocated once, captured in a CUDA graph and
replayed per epoch.  Every node of the graph is one of the library's own kernels; the programmatic-dependent-launch edges
between them are kept by the capture.  The model's parameters are updated in place, so the usual objects (``model``,
``model.state_dict()``) stay valid; ``logits`` / ``loss`` are static tensors holding the last epoch's pre-step output, which is
what the reference's loop reads after the step (``train.py:206-226``).

The autograd path (``model(g, x)`` ... ``loss.backward()`` ... ``FusedAdam.step()``) remains what the UNCHANGED driver uses;
both run the same kernels in the same order and are compared in ``tests/test_gpu_epoch.py``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops
from .loss import _class_weights
from .nn import GNN32, _ptr_array


class TrainStep:
    def __init__(self, model: GNN32, g, features: torch.Tensor, labels: torch.Tensor, train_index, i_weight,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, use_graph: bool = True, warmup: int = 2):
        lib = _lib.load()
        self.lib, self.model, self.g = lib, model, g
        self.params = [p for p in model.hot_path_parameters()]
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.PlagnnError("TrainStep needs contiguous float32 CUDA parameters (no CPU fallback)")
        dev = self.params[0].device
        self.device = dev
        with torch.cuda.device(dev):
            self.state = model._engine_state(g)
            self.xa = ops.aligned(features.detach())
            self.labels = labels.detach() if labels.stride(1) == 1 else labels.detach().contiguous()
            ops._require_cuda_f32(self.xa, self.labels)
            if not isinstance(train_index, torch.Tensor):
                train_index = torch.as_tensor(np.asarray(train_index), dtype=torch.int64)
            self.index = train_index.to(device=dev, dtype=torch.int64).contiguous()
            self.cw, self.cwp1 = _class_weights(i_weight, dev)
            n, c = self.state.shape.num_nodes, self.state.shape.classes
            self.n, self.c = n, c
            self.arena = torch.zeros(self.state.arena_bytes, dtype=torch.uint8, device=dev)
            self.logits = ops.alloc(n, c, dev, zero=True)
            self.dprob = ops.alloc(n, c, dev, zero=True)
            self.loss = torch.zeros(1, device=dev)
            self.grads = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
            self.exp_avg = [torch.zeros_like(g_) for g_ in self.grads]
            self.exp_avg_sq = [torch.zeros_like(g_) for g_ in self.grads]
            rows = [(p.data_ptr(), g_.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
                    for p, g_, m, v in zip(self.params, self.grads, self.exp_avg, self.exp_avg_sq)]
            self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
            self.max_numel = max(p.numel() for p in self.params)
            self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
            self.scalars = torch.zeros(4, device=dev)
            self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
            self.bce_bytes = lib.plagnn_bce_workspace_bytes(self.index.numel(), c)
            self.bce_ws = torch.empty(max(self.bce_bytes, 16), dtype=torch.uint8, device=dev)
            self._p_ptrs = _ptr_array([p.detach() for p in self.params])
            self._g_ptrs = _ptr_array(self.grads)
            self.graph = None
            self.epochs = 0
            self.launches_per_epoch = None
            if use_graph:
                self._capture(warmup)

    # one epoch, enqueued on `stream` (a raw cudaStream_t value)
    def _enqueue(self, stream: int) -> None:
        lib, st, sh = self.lib, ctypes.c_void_p(stream), self.state.shape_ref
        xa, lg, dp = self.xa, self.logits, self.dprob
        _lib.check(lib.plagnn_gnn32_forward(sh, xa.data_ptr(), xa.stride(0), self._p_ptrs, self.arena.data_ptr(), self.arena.numel(),
                                            lg.data_ptr(), lg.stride(0), st), "gnn32_forward")
        _lib.check(lib.plagnn_bce_weighted(lg.data_ptr(), lg.stride(0), self.labels.data_ptr(), self.labels.stride(0),
                                           self.index.data_ptr(), self.index.numel(), self.n, self.c, self.cw.data_ptr(),
                                           self.cwp1.data_ptr(), 1.0, self.loss.data_ptr(), dp.data_ptr(), dp.stride(0),
                                           self.bce_ws.data_ptr(), self.bce_ws.numel(), st), "bce_weighted")
        _lib.check(lib.plagnn_gnn32_backward(sh, xa.data_ptr(), xa.stride(0), self._p_ptrs, self.arena.data_ptr(), self.arena.numel(),
                                             lg.data_ptr(), lg.stride(0), dp.data_ptr(), dp.stride(0), self._g_ptrs, None, 0, st),
                   "gnn32_backward")
        _lib.check(lib.plagnn_adam_multi_devstep(self.table.data_ptr(), len(self.params), self.max_numel, self.lr, self.betas[0],
                                                 self.betas[1], self.eps, self.step_count.data_ptr(), self.scalars.data_ptr(), st),
                   "adam_multi_devstep")

    def _capture(self, warmup: int) -> None:
        """Warm-up epochs before the capture (module loading, function attributes, tensor-map encoding all happen on a first
        launch) run on the capture stream and leave no trace: parameters, Adam state and the step count are restored."""
        with torch.cuda.device(self.device):
            saved = [p.detach().clone() for p in self.params]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            before = ops.launch_count()
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):
                    self._enqueue(side.cuda_stream)
            self.launches_per_epoch = (ops.launch_count() - before) // max(warmup, 1)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                for p, s_ in zip(self.params, saved):
                    p.copy_(s_)
                for t in self.exp_avg + self.exp_avg_sq:
                    t.zero_()
                self.step_count.zero_()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self._enqueue(torch.cuda.current_stream().cuda_stream)
            self.graph = graph

    def step(self) -> None:
        """One epoch (forward, loss on the training rows, backward, Adam step) on the current stream."""
        with torch.cuda.device(self.device):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._enqueue(torch.cuda.current_stream().cuda_stream)
        self.epochs += 1

    def run(self, epochs: int) -> None:
        for _ in range(epochs):
            self.step()
