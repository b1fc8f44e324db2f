"""Graph object of the drop-in: the minimal surface the reference touches on a DGLGraph
(``code/utils.py:44-49``, ``code/main_normal.py:66``, ``code/train.py:145-146,179``) plus the
device-side in-edge / out-edge CSR that the aggregation kernels walk.

Edge direction and numbering follow DGL: edge e goes ``src[e] -> dst[e]``, its id is its position in
the COO arrays, ``add_self_loop`` appends (i,i) for i in 0..N-1 after them.  The CSR arrays are int32 on
the device (DGL would hold int64 ids; values are identical).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import check

# neighbours per aggregation work item (one warp).  Swept on B200, PPI-shaped graph, F = 503 (tools/chunk_sweep.py):
# 64 -> 0.304 ms, 128 -> 0.248, 256 -> 0.223, 512 -> 0.212, 1024 -> 0.311 (hub tail), no split -> 3.77 ms.
DEFAULT_CHUNK = 512


class Csr:
    """One direction of the adjacency on the device: rows = `key` endpoint, entries = other endpoint."""

    def __init__(self, indptr, indices, eids, num_rows, num_edges, chunk=DEFAULT_CHUNK):
        self.indptr, self.indices, self.eids = indptr, indices, eids
        self.num_rows, self.num_edges = int(num_rows), int(num_edges)
        self.chunk = int(chunk)
        self.plan = None
        self.counts = [0, 0, 0]
        self.counts_c = (ctypes.c_int64 * 3)()
        self._build_plan()

    def _build_plan(self):
        lib = _lib.load()
        dev = self.indptr.device
        nb = lib.plagnn_spmm_plan_bytes(self.num_rows, self.num_edges, self.chunk)
        self.plan = torch.empty(nb, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.plagnn_spmm_plan_build(ops._p(self.indptr), self.num_rows, self.num_edges, self.chunk,
                                             ops._p(self.plan), nb, self.counts_c, ops._stream()), "spmm_plan_build")
        self.counts = [int(self.counts_c[i]) for i in range(3)]

    @property
    def degrees(self) -> torch.Tensor:
        return (self.indptr[1:] - self.indptr[:-1])


def build_csr(key: torch.Tensor, other: torch.Tensor, num_nodes: int, add_self_loop: bool,
              chunk: int = DEFAULT_CHUNK, num_other: int = 0) -> Csr:
    """Stable sort of edge ids by `key` on the device (plagnn_csr_build).  key/other: int32 CUDA tensors.
    num_nodes = id range of `key` (= number of rows); num_other = id range of `other` (0: same)."""
    lib = _lib.load()
    if not key.is_cuda:
        raise _lib.PlagnnError("build_csr needs CUDA tensors (no CPU fallback)")
    key = key.to(torch.int32).contiguous()
    other = other.to(torch.int32).contiguous()
    e = key.numel()
    n = int(num_nodes)
    ep = e + (n if add_self_loop else 0)
    dev = key.device
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(ep, 1), dtype=torch.int32, device=dev)
    eids = torch.empty(max(ep, 1), dtype=torch.int32, device=dev)
    nb = lib.plagnn_csr_build_workspace_bytes(n, e, int(add_self_loop))
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.plagnn_csr_build(ops._p(key), ops._p(other), e, n, int(num_other), int(add_self_loop), ops._p(indptr),
                                   ops._p(indices), ops._p(eids), ops._p(ws), nb, ops._stream()), "csr_build")
    del ws
    return Csr(indptr, indices[:ep], eids[:ep], n, ep, chunk)


class _NodeSpace:
    def __init__(self, graph, ids):
        self._g, self._ids = graph, ids

    @property
    def data(self):
        return _NodeData(self._g, self._ids)


class _NodeData:
    def __init__(self, graph, ids):
        self._g, self._ids = graph, ids

    def __setitem__(self, key, value):
        g = self._g
        ids = self._ids
        value = torch.as_tensor(value)
        full = isinstance(ids, (list, range)) and len(ids) == g.num_nodes() and (
            isinstance(ids, range) or ids == list(range(g.num_nodes())))
        if full or (isinstance(ids, slice) and ids == slice(None)):
            g.ndata[key] = value
            return
        if key not in g.ndata:
            g.ndata[key] = torch.zeros((g.num_nodes(),) + tuple(value.shape[1:]), dtype=value.dtype)
        g.ndata[key][torch.as_tensor(ids, dtype=torch.long)] = value

    def __getitem__(self, key):
        return self._g.ndata[key][torch.as_tensor(self._ids, dtype=torch.long)]


class _NodeView:
    def __init__(self, graph):
        self._g = graph

    def __getitem__(self, ids):
        return _NodeSpace(self._g, ids)

    def __call__(self):
        return torch.arange(self._g.num_nodes())


class Graph:
    """COO graph + node data; device CSR/CSC built on first use after ``.to('cuda')``."""

    def __init__(self, src, dst, num_nodes, self_loops: bool = False, chunk: int = DEFAULT_CHUNK):
        self._src = torch.as_tensor(np.asarray(src) if not isinstance(src, torch.Tensor) else src).to(torch.int32)
        self._dst = torch.as_tensor(np.asarray(dst) if not isinstance(dst, torch.Tensor) else dst).to(torch.int32)
        if self._src.shape != self._dst.shape or self._src.dim() != 1:
            raise ValueError("src and dst must be 1-D arrays of equal length")
        self._n = int(num_nodes)
        if self._src.numel() and self._n <= int(max(self._src.max(), self._dst.max())):
            raise ValueError("num_nodes is smaller than the largest node id + 1")
        self._loops = bool(self_loops)
        self.chunk = chunk
        self.ndata = {}
        self.edata = {}
        self._csc = None
        self._csr = None

    # ---- the DGLGraph surface the reference uses --------------------------------------------------
    @property
    def nodes(self):
        return _NodeView(self)

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return self._src.numel() + (self._n if self._loops else 0)

    number_of_edges = num_edges

    @property
    def device(self):
        return self._src.device

    def edges(self):
        """(src, dst) int64 tensors including appended self-loops, in edge-id order."""
        s, d = self._src.long(), self._dst.long()
        if self._loops:
            loop = torch.arange(self._n, device=s.device)
            s, d = torch.cat([s, loop]), torch.cat([d, loop])
        return s, d

    def to(self, device):
        device = torch.device(device)
        g = Graph.__new__(Graph)
        g._src, g._dst = self._src.to(device), self._dst.to(device)
        g._n, g._loops, g.chunk = self._n, self._loops, self.chunk
        g._csc = g._csr = None
        g.ndata, g.edata = {}, {}
        for k, v in self.ndata.items():
            g.ndata[k] = _to_device_aligned(v, device)
        for k, v in self.edata.items():
            g.edata[k] = v.to(device)
        return g

    def add_self_loop(self):
        if self._loops:   # second application: materialise the first set, then append again
            s, d = self.edges()
            g = Graph(s, d, self._n, True, self.chunk)
        else:
            g = Graph(self._src, self._dst, self._n, True, self.chunk)
        g.ndata = dict(self.ndata)
        return g

    # ---- device structure --------------------------------------------------------------------------
    def csc(self) -> Csr:
        """In-edge CSR (rows = destination, entries = source): what update_all(copy_u, reduce) walks."""
        if self._csc is None:
            self._csc = build_csr(self._dst, self._src, self._n, self._loops, self.chunk)
        return self._csc

    def csr(self) -> Csr:
        """Out-edge CSR (rows = source, entries = destination): the transposed aggregation."""
        if self._csr is None:
            self._csr = build_csr(self._src, self._dst, self._n, self._loops, self.chunk)
        return self._csr

    def in_degrees(self) -> torch.Tensor:
        return self.csc().degrees

    def has_duplicate_edges(self) -> bool:
        """True if some (src, dst) pair occurs more than once (bookkeeping, cached; torch ops on the COO arrays)."""
        if getattr(self, "_dup", None) is None:
            key = self._src.long() * self._n + self._dst.long()
            dup = torch.unique(key).numel() != key.numel()
            if self._loops and not dup:
                dup = bool((self._src == self._dst).any())
            self._dup = bool(dup)
        return self._dup


def _to_device_aligned(v: torch.Tensor, device) -> torch.Tensor:
    """2-D float32 node data lands in a row-padded buffer so that kernels can use it in place."""
    if device.type == "cuda" and v.dim() == 2 and v.dtype == torch.float32:
        out = ops.alloc(v.shape[0], v.shape[1], device, zero=True)
        out.copy_(v, non_blocking=True)
        return out
    return v.to(device)


def graph(data, num_nodes=None, idtype=None, device=None) -> Graph:
    """dgl.graph((src, dst), num_nodes=N)  (code/utils.py:44)."""
    src, dst = data
    src = np.asarray(src) if not isinstance(src, torch.Tensor) else src
    dst = np.asarray(dst) if not isinstance(dst, torch.Tensor) else dst
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1 if len(src) else 0
    g = Graph(src, dst, num_nodes)
    return g.to(device) if device is not None else g


def add_self_loop(g: Graph) -> Graph:
    """dgl.add_self_loop(g)  (code/utils.py:45)."""
    return g.add_self_loop()
