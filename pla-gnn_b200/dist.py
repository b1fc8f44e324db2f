"""1-D row partition of the aggregation across GPUs (BASELINE.json configs[3]).

The reference has no distributed code at all (single process, one ``-d`` device string,
``code/main_normal.py:30,66``); the PPI graph fits one GPU, so partitioning is used only for the scaled
synthetic graph.  One process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch) for the exchange:

* destination rows are split into ``world`` equal contiguous blocks (node ids are shuffled by the generator, so
  rows and in-edges balance together); rank p owns ``x[rows_p]``, the in-edge CSR of its rows with source ids in
  the gathered numbering, its loss rows and a full replica of the weights;
* forward, per layer:  t_p = h_p W^T (local GEMM)  ->  all-gather t  ->  out_p = act(scale * A_p t + b) (local SpMM);
* backward, per layer: dt = A_p^T (scale * dz_p) over ALL source rows (local transposed SpMM)  ->  reduce-scatter
  ->  dt_p;  dW = dt_p^T h_p  ->  one all-reduce per step with the other weight gradients.

Overlap.  Every rank's block is cut into ``chunks`` row chunks and gathered matrices are laid out CHUNK-MAJOR
(gathered row = chunk * world * cr + rank * cr + row_in_chunk), so the all-gather of local chunk c fills one
contiguous slab and the reduce-scatter of slab c yields local chunk c.  Forward: while chunk c of layer l+1 is
being gathered, the aggregation (layer l) and projection (layer l+1) of chunk c+1 run on the compute stream
(row-range SpMM launches, ``plagnn_spmm_sum_rows``).  Backward: the transposed aggregation produces slab c+1 while
slab c is being reduce-scattered.  Only the first layer's gather and the last chunk of each exchange stay exposed.

Halo pruning (sending only referenced rows) is pointless here: at ~100 in-edges per node on a power-law graph
every rank references essentially every source row, so the exchange is a plain all-gather; fetching neighbour
rows from peer memory inside the SpMM would move each remote row E/(N*world) ~ 12x instead of once.

The choreography is separated from the compute backend so that the N > 1 path runs on CPU with gloo in the tests
(they inject a CPU backend with the same method names; the product uses the CUDA kernels + NCCL).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def block_bounds(num_nodes: int, world: int):
    """Equal contiguous row blocks: (rows_per_rank, [r0_0, ..., r0_world]); the last block may be short."""
    per = (num_nodes + world - 1) // world
    return per, [min(p * per, num_nodes) for p in range(world + 1)]


class RowPartitionPlan:
    """Pure index arithmetic (any device, no kernels): which edges a rank owns and how they are renumbered."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int, rank: int, world: int, chunks: int = 1):
        self.num_nodes, self.rank, self.world, self.chunks = int(num_nodes), int(rank), int(world), int(chunks)
        self.per_raw, self.bounds = block_bounds(num_nodes, world)
        self.cr = (self.per_raw + chunks - 1) // chunks       # rows per chunk
        self.per = self.cr * chunks                           # padded rows per rank
        self.r0, self.r1 = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.r1 - self.r0
        self.n_padded = self.per * world                      # rows of a gathered matrix
        mask = (dst >= self.r0) & (dst < self.r1)
        self.edge_ids = torch.nonzero(mask, as_tuple=False).flatten()   # positions in the global COO arrays
        self.src_gathered = self.gathered_id(src[mask])
        self.dst_local = dst[mask] - self.r0                  # ids in [0, n_local)
        in_deg = torch.bincount(self.dst_local, minlength=self.per).to(torch.float32)
        self.scale_local = 1.0 / in_deg.clamp(min=1.0)       # right normalisation (mean over in-edges)

    def gathered_id(self, g: torch.Tensor) -> torch.Tensor:
        """Global node id -> row of a chunk-major gathered matrix."""
        owner = g // self.per_raw
        i = g - owner * self.per_raw
        return (i // self.cr) * (self.world * self.cr) + owner * self.cr + (i % self.cr)

    def local_rows(self, c: int):
        return c * self.cr, (c + 1) * self.cr

    def slab_rows(self, c: int):
        return c * self.world * self.cr, (c + 1) * self.world * self.cr

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
    """Device structures of one rank: in-edge CSR of the owned rows (gathered source ids) and its transpose."""

    def __init__(self, plan: RowPartitionPlan, weight_global: torch.Tensor | None, build_csr, device):
        self.plan = plan
        s = plan.src_gathered.to(device=device, dtype=torch.int32)
        d = plan.dst_local.to(device=device, dtype=torch.int32)
        # rows = padded local rows (per), entries = rows of the gathered matrix
        self.csc = build_csr(d, s, plan.per, False, num_other=plan.n_padded)
        # transpose: rows = gathered (source) rows, entries = local destination rows
        self.csr_t = build_csr(s, d, plan.n_padded, False, num_other=plan.per)
        self.edge_weight = None if weight_global is None else weight_global[plan.edge_ids.to(weight_global.device)].to(device)
        self.scale = plan.scale_local.to(device)


def dist_gcn_forward_backward(model: DistGCN, pg, h0_local, backend, group=None, loss_grad_fn=None, act_leaky=True):
    """One forward + backward of the partitioned GCN.  h0_local: [per, F0] (rows past n_local are zero).
    Returns (out_local, grads) with grads = [W0, b0, W1, b1, ...] already summed over ranks."""
    plan = pg.plan
    world, chunks = plan.world, plan.chunks
    n_layers = len(model.weights)
    dims = model.dims
    hs = [h0_local]                  # layer inputs (local rows)
    outs = []                        # layer outputs (local rows)
    t_full_prev = None
    # ---------------- forward ----------------
    for li in range(n_layers):
        w = model.weights[li].detach()
        t_local = backend.alloc(plan.per, dims[li + 1])
        t_full = backend.alloc(plan.n_padded, dims[li + 1])
        if li > 0:
            out_prev = backend.alloc(plan.per, dims[li])
        pending = []
        for c in range(chunks):
            a, b = plan.local_rows(c)
            if li > 0:   # aggregation of the previous layer for this row chunk, then this layer's projection
                backend.spmm_rows(pg.csc, ("csc", c), a, b, t_full_prev, out_prev, pg.edge_weight, pg.scale,
                                  model.biases[li - 1].detach(), act=act_leaky)
                h_c = backend.rows(out_prev, a, b)
            else:
                h_c = backend.rows(h0_local, a, b)
            backend.gemm_nt_into(h_c, w, backend.rows(t_local, a, b))
            sa, sb = plan.slab_rows(c)
            pending.append(backend.all_gather_chunk(backend.rows(t_full, sa, sb), backend.rows(t_local, a, b), world, group))
        backend.wait(pending)
        if li > 0:
            outs.append(out_prev)
            hs.append(out_prev)
        t_full_prev = t_full
    out_last = backend.alloc(plan.per, dims[-1])
    backend.spmm_rows(pg.csc, ("csc", -1), 0, plan.per, t_full_prev, out_last, pg.edge_weight, pg.scale,
                      model.biases[-1].detach(), act=False)
    outs.append(out_last)
    # ---------------- backward ----------------
    d_out = loss_grad_fn(out_last) if loss_grad_fn is not None else torch.ones_like(out_last)
    grads = [None] * (2 * n_layers)
    dz = d_out
    for li in reversed(range(n_layers)):
        last = li + 1 == n_layers
        w = model.weights[li].detach()
        dzb = backend.act_backward(dz, outs[li] if (not last and act_leaky) else None, None)
        grads[2 * li + 1] = backend.colsum(dzb)
        dzs = backend.act_backward(dzb, None, pg.scale)
        dt_partial = backend.alloc(plan.n_padded, dims[li + 1])
        dt_local = backend.alloc(plan.per, dims[li + 1])
        pending = []
        for c in range(chunks):
            sa, sb = plan.slab_rows(c)
            backend.spmm_rows(pg.csr_t, ("csr_t", c), sa, sb, dzs, dt_partial, pg.edge_weight, None, None, act=False)
            a, b = plan.local_rows(c)
            pending.append(backend.reduce_scatter_chunk(backend.rows(dt_local, a, b), backend.rows(dt_partial, sa, sb),
                                                        world, group))
        backend.wait(pending)
        grads[2 * li] = backend.gemm_tn(dt_local, hs[li])                                  # [O, F]
        if li > 0:
            dz = backend.gemm_nn(dt_local, w)                                              # [per, F]
    return out_last, backend.all_reduce_grads(grads, world, group)


class CudaBackend:
    """Compute + collectives on the CUDA kernels / NCCL.  (Tests use a CPU twin with the same method names.)"""

    def __init__(self, pg: PartitionedGraph):
        from . import ops
        self.ops = ops
        plan = pg.plan
        self.ranges = {}
        for c in range(plan.chunks):
            self.ranges[("csc", c)] = ops.plan_range(pg.csc, *plan.local_rows(c))
            self.ranges[("csr_t", c)] = ops.plan_range(pg.csr_t, *plan.slab_rows(c))
        self.ranges[("csc", -1)] = ops.plan_range(pg.csc, 0, plan.per)

    def alloc(self, rows, cols):
        return self.ops.alloc(rows, cols, torch.cuda.current_device())

    @staticmethod
    def rows(t, a, b):
        return t[a:b]

    @staticmethod
    def _padded(t):
        """The row-padded storage behind a [rows, cols] view (contiguous: what NCCL gets)."""
        return t.as_strided((t.shape[0], t.stride(0)), (t.stride(0), 1), t.storage_offset())

    def gemm_nt_into(self, a, w, out):
        ops = self.ops
        a = ops.aligned(a)
        ops.gemm(a.shape[0], w.shape[0], [(a, 0, ops.aligned(w), 0, a.shape[1])], out=out)

    def gemm_nn(self, a, w):
        ops = self.ops
        a = ops.aligned(a)
        return ops.gemm(a.shape[0], w.shape[1], [(a, 0, ops.aligned(w), 1, a.shape[1])])

    def gemm_tn(self, a, b):
        ops = self.ops
        a, b = ops.aligned(a), ops.aligned(b)
        out = torch.empty((a.shape[1], b.shape[1]), device=a.device, dtype=torch.float32)
        return ops.gemm(a.shape[1], b.shape[1], [(a, 1, b, 1, a.shape[0])], out=out)

    def spmm_rows(self, csx, key, a, b, x, out, w, scale, bias, act):
        ops = self.ops
        ops.spmm_sum_rows(csx, self.ranges[key], x, out, w=w, scale=scale, bias=bias,
                          act=ops.ACT_LEAKY if act else ops.ACT_NONE)

    def act_backward(self, dy, y, scale):
        ops = self.ops
        if y is None and scale is None:
            return dy
        return ops.act_backward(dy, y, ops.ACT_LEAKY if y is not None else ops.ACT_NONE, row_scale=scale)

    def colsum(self, x):
        return self.ops.colsum(self.ops.aligned(x))

    def all_gather_chunk(self, slab, local_chunk, world, group):
        if world == 1:
            slab.copy_(local_chunk)
            return None
        return dist.all_gather_into_tensor(self._padded(slab), self._padded(local_chunk), group=group, async_op=True)

    def reduce_scatter_chunk(self, out_chunk, slab, world, group):
        if world == 1:
            out_chunk.copy_(slab)
            return None
        return dist.reduce_scatter_tensor(self._padded(out_chunk), self._padded(slab), op=dist.ReduceOp.SUM, group=group,
                                          async_op=True)

    @staticmethod
    def wait(pending):
        for h in pending:
            if h is not None:
                h.wait()              # the compute stream waits for the NCCL stream; the host does not block

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
