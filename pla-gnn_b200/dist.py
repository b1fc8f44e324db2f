"""1-D row partition of the aggregation across GPUs (BASELINE.json configs[3]).

The reference has no distributed code at all (single process, one ``-d`` device string,
``code/main_normal.py:30,66``); the PPI graph fits one GPU, so partitioning is used only for the scaled
synthetic graph.  One process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch) for the exchange:

* destination rows are split into ``world`` equal contiguous blocks (node ids are shuffled by the generator, so
  rows and in-edges balance together); rank p owns ``x[rows_p]``, the in-edge CSR of its rows with GLOBAL
  source ids, its labels / loss rows and a full replica of the weights;
* forward, per layer:  t_p = h_p W^T (local GEMM)  ->  all-gather t  ->  out_p = act(scale * A_p t + b) (local SpMM);
* backward, per layer: dt = A_p^T (scale * dz_p) over ALL source rows (local transposed SpMM)  ->  reduce-scatter
  ->  dt_p;  dW = dt_p^T h_p  ->  all-reduce once per step with the other weight gradients.

Halo pruning (sending only referenced rows) is pointless here: at ~100 in-edges per node on a power-law graph
every rank references essentially every source row, so the exchange is a plain all-gather.  Fetching neighbour
rows directly from peer memory inside the SpMM would move each remote row E/(N*world) ~ 12x instead of once.

The communication choreography is separated from the compute backend (``ops``-like object) so that the N > 1 path
can be exercised on CPU with the gloo backend (tests inject a CPU backend; the product uses the CUDA kernels).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def block_bounds(num_nodes: int, world: int):
    """Equal contiguous row blocks: (rows_per_rank, [r0_0, ..., r0_world]); the last block may be short."""
    per = (num_nodes + world - 1) // world
    return per, [min(p * per, num_nodes) for p in range(world + 1)]


def to_padded_ids(ids: torch.Tensor, per: int) -> torch.Tensor:
    """Global node id -> row in the all-gathered, per-rank padded matrix [world * per, F] (identity when the blocks
    are full: rank = id // per, local = id % per -> rank*per + local == id).  Kept explicit for clarity."""
    return ids


class RowPartitionPlan:
    """Pure index arithmetic (any device, no kernels): which edges a rank owns and how they are renumbered."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int, rank: int, world: int):
        self.num_nodes, self.rank, self.world = int(num_nodes), int(rank), int(world)
        self.per, self.bounds = block_bounds(num_nodes, world)
        self.r0, self.r1 = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.r1 - self.r0
        self.n_padded = self.per * world                      # rows of an all-gathered matrix
        mask = (dst >= self.r0) & (dst < self.r1)
        self.edge_ids = torch.nonzero(mask, as_tuple=False).flatten()   # positions in the global COO arrays
        self.src_global = src[mask]                           # ids in [0, N) == rows of the gathered matrix
        self.dst_local = dst[mask] - self.r0                  # ids in [0, n_local)
        in_deg = torch.bincount(self.dst_local, minlength=self.per).to(torch.float32)
        self.scale_local = 1.0 / in_deg.clamp(min=1.0)       # right normalisation (mean over in-edges)

    @property
    def num_local_edges(self):
        return int(self.edge_ids.numel())


class DistGCN(torch.nn.Module):
    """L-layer weighted-sum GCN over a row partition.  Parameters are replicated (same seed on every rank)."""

    def __init__(self, dims, seed: int = 0):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)
        self.weights = torch.nn.ParameterList()
        self.biases = torch.nn.ParameterList()
        for i in range(len(dims) - 1):
            bound = 1.0 / math.sqrt(dims[i])
            self.weights.append(torch.nn.Parameter((torch.rand(dims[i + 1], dims[i], generator=gen) * 2 - 1) * bound))
            self.biases.append(torch.nn.Parameter(torch.zeros(dims[i + 1])))
        self.dims = list(dims)


class PartitionedGraph:
    """Device structures of one rank: in-edge CSR of the owned rows (global sources) and its transpose."""

    def __init__(self, plan: RowPartitionPlan, weight_global: torch.Tensor | None, build_csr, device):
        self.plan = plan
        s = plan.src_global.to(device=device, dtype=torch.int32)
        d = plan.dst_local.to(device=device, dtype=torch.int32)
        # rows = padded local rows (per), entries = global source ids in the gathered matrix
        self.csc = build_csr(d, s, plan.per, False, num_other=plan.n_padded)
        # transpose: rows = all (padded) source rows, entries = local destination rows
        self.csr_t = build_csr(s, d, plan.n_padded, False, num_other=plan.per)
        self.edge_weight = None if weight_global is None else weight_global[plan.edge_ids.to(weight_global.device)].to(device)
        self.scale = plan.scale_local.to(device)


def dist_gcn_forward_backward(model: DistGCN, pg, h0_local, backend, group=None, loss_grad_fn=None, act_leaky=True):
    """One forward + backward of the partitioned GCN.  h0_local: [per, F0] (rows past n_local are zero).
    backend provides gemm_nt(a, w) = a @ w.T, gemm_nn(a, w) = a @ w, gemm_tn(a, b) = a.T @ b,
    spmm(csx, x, w, scale, bias, act) and act_backward(dy, y, scale).  Returns (out_local, grads list) where grads
    follows [W0, b0, W1, b1, ...] and is already all-reduced (sum over ranks)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_layers = len(model.weights)
    hs, outs = [h0_local], []
    h = h0_local
    for li in range(n_layers):
        w, b = model.weights[li].detach(), model.biases[li].detach()
        t_local = backend.gemm_nt(h, w)                                   # [per, O]
        t_full = backend.all_gather_rows(t_local, world, group)           # [world*per, O]
        last = li + 1 == n_layers
        out = backend.spmm(pg.csc, t_full, pg.edge_weight, pg.scale, b, act=(not last) and act_leaky)
        outs.append(out)
        h = out
        hs.append(h)
    d_out = loss_grad_fn(outs[-1]) if loss_grad_fn is not None else torch.ones_like(outs[-1])
    grads = [None] * (2 * n_layers)
    dz = d_out
    for li in reversed(range(n_layers)):
        last = li + 1 == n_layers
        w = model.weights[li].detach()
        # dz = d out * act'(out); bias gradient before the row scaling; then scale rows for the transposed SpMM
        dzb = backend.act_backward(dz, outs[li] if not last and act_leaky else None, None)
        grads[2 * li + 1] = backend.colsum(dzb)
        dzs = backend.act_backward(dzb, None, pg.scale)
        dt_partial = backend.spmm(pg.csr_t, dzs, pg.edge_weight, None, None, act=False)     # [world*per, O]
        dt_local = backend.reduce_scatter_rows(dt_partial, world, group)                  # [per, O]
        grads[2 * li] = backend.gemm_tn(dt_local, hs[li])                                  # [O, F]
        if li > 0:
            dz = backend.gemm_nn(dt_local, w)                                              # [per, F]
    flat = backend.all_reduce_grads(grads, world, group)
    return outs[-1], flat


class CudaBackend:
    """Compute + collectives on the CUDA kernels / NCCL.  (Tests use a CPU twin with the same method names.)"""

    def __init__(self):
        from . import ops
        self.ops = ops

    def gemm_nt(self, a, w):
        ops = self.ops
        a = ops.aligned(a)
        return ops.gemm(a.shape[0], w.shape[0], [(a, 0, ops.aligned(w), 0, a.shape[1])])

    def gemm_nn(self, a, w):
        ops = self.ops
        a = ops.aligned(a)
        return ops.gemm(a.shape[0], w.shape[1], [(a, 0, ops.transpose(w), 0, a.shape[1])])

    def gemm_tn(self, a, b):
        ops = self.ops
        a, b = ops.aligned(a), ops.aligned(b)
        out = torch.empty((a.shape[1], b.shape[1]), device=a.device, dtype=torch.float32)
        return ops.gemm(a.shape[1], b.shape[1], [(a, 1, b, 1, a.shape[0])], out=out)

    def spmm(self, csx, x, w, scale, bias, act):
        ops = self.ops
        return ops.spmm_sum(csx, x, w=w, scale=scale, bias=bias, act=ops.ACT_LEAKY if act else ops.ACT_NONE)

    def act_backward(self, dy, y, scale):
        ops = self.ops
        if y is None and scale is None:
            return dy
        return ops.act_backward(dy, y, ops.ACT_LEAKY if y is not None else ops.ACT_NONE, row_scale=scale)

    def colsum(self, x):
        return self.ops.colsum(self.ops.aligned(x))

    def all_gather_rows(self, t_local, world, group):
        if world == 1:
            return t_local
        # gather the padded buffers (row pitch included) so that the result is directly a row-aligned matrix
        base = t_local._base if t_local._base is not None else t_local
        full = torch.empty((world * base.shape[0], base.shape[1]), device=base.device, dtype=base.dtype)
        dist.all_gather_into_tensor(full, base.contiguous(), group=group)
        return full[:, :t_local.shape[1]]

    def reduce_scatter_rows(self, partial, world, group):
        if world == 1:
            return partial
        base = partial._base if partial._base is not None else partial
        out = torch.empty((base.shape[0] // world, base.shape[1]), device=base.device, dtype=base.dtype)
        dist.reduce_scatter_tensor(out, base.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out[:, :partial.shape[1]]

    def all_reduce_grads(self, grads, world, group):
        if world == 1:
            return grads
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out, off = [], 0
        for g in grads:
            out.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return out
