"""Partitioned aggregation across GPUs (BASELINE.json configs[3]).

The reference has no distributed code at all (single process, one ``-d`` device string,
``code/main_normal.py:30,66``); the PPI graph fits one GPU, so partitioning is used only for the scaled
synthetic graph.  One process per GPU; the exchange steps go through the C ABI's NCCL wrappers
(``plagnn_nccl_*``, NVLink / NVSwitch) on the GPU and through ``torch.distributed`` (gloo) in the CPU tests.

Two ways to split the aggregation, two reducers each (weighted sum = the north_star's ``u_mul_e`` + ``sum`` with right
normalisation; max = the reference's SAGEConv 'pool'):

**Row partition** (``mode="rows"``, north_star: "1D row partitioning with a halo feature all-gather"):
  destination rows are split into ``world`` contiguous blocks (equal, or balanced by in-edge count); rank p owns
  ``x[rows_p]``, the in-edge CSR of its rows with source ids in the gathered numbering, its loss rows and a replica of the
  weights.  Per layer: local GEMM -> all-gather of the projected rows -> local SpMM; backward: local transposed SpMM (or
  arg-scatter) over ALL source rows -> reduce-scatter.  Every rank receives the whole N x F matrix per layer and direction.
  Overlap: a rank's block is cut into ``chunks`` row chunks, gathered matrices are CHUNK-MAJOR (gathered row = chunk *
  world * cr + rank * cr + row_in_chunk), so the all-gather of local chunk c fills one contiguous slab and the
  reduce-scatter of slab c yields local chunk c; the aggregation / projection of chunk c+1 run while chunk c is in flight.

**Feature partition** (``mode="cols"``): the aggregation is independent per feature column, so rank p aggregates ALL rows
  but only columns [p F/P, (p+1) F/P) over a replica of the graph structure; the dense products stay row-partitioned.
  Around every aggregation one all-to-all turns "my rows x all columns" into "all rows x my columns" and one turns it
  back.  Each moves N F (P-1)/P^2 floats per GPU — P times less than the all-gather (112 MB instead of 0.9 GB at P = 8,
  F = 256) — the max reducer's backward needs no reduction across ranks at all (a rank holds every row of its columns),
  and at P = 8 the N x 32 slice a rank gathers from (128 MB) is L2-sized.  Costs: the structure is replicated (0.8 GB
  per direction at 100 M edges — nothing against 180 GB) and narrow rows (128 bytes at P = 8) need the sub-warp kernel.

Halo pruning (sending only referenced rows) is pointless here: at ~100 in-edges per node on a power-law graph every rank
references essentially every source row; fetching neighbour rows from peer memory inside the SpMM would move each remote
row E/(N*world) ~ 12x instead of once.

The choreography is separated from the compute backend so that the N > 1 paths run on CPU with gloo in the tests
(they inject a CPU backend with the same method names; the product uses the CUDA kernels + NCCL).
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.distributed as dist


def block_bounds(num_nodes: int, world: int):
    """Equal contiguous row blocks: (rows_per_rank, [r0_0, ..., r0_world]); the last block may be short."""
    per = (num_nodes + world - 1) // world
    return per, [min(p * per, num_nodes) for p in range(world + 1)]


def block_bounds_by_edges(in_degree: torch.Tensor, world: int):
    """Contiguous row blocks balanced by in-edge count (power-law => balance edges, not nodes):
    (largest block, [r0_0, ..., r0_world]).  Block p ends at the first row where the running in-edge count reaches
    (p + 1) E / world."""
    n = int(in_degree.numel())
    csum = torch.cumsum(in_degree.to(torch.int64), 0)
    total = int(csum[-1]) if n else 0
    targets = torch.tensor([(p * total + world - 1) // world for p in range(1, world)], dtype=torch.int64, device=csum.device)
    cuts = (torch.searchsorted(csum, targets, right=False) + 1).clamp(max=n).tolist() if world > 1 else []
    bounds = [0] + cuts + [n]
    for i in range(1, len(bounds)):                       # monotone even for degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    per = max(bounds[i + 1] - bounds[i] for i in range(world))
    return per, bounds


class RowPartitionPlan:
    """Pure index arithmetic (any device, no kernels): which edges a rank owns and how they are renumbered.
    balance = "rows" (equal blocks; node ids are shuffled by the generator, so rows and in-edges balance together)
    or "edges" (blocks of equal in-edge count)."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int, rank: int, world: int, chunks: int = 1,
                 balance: str = "rows"):
        self.num_nodes, self.rank, self.world, self.chunks = int(num_nodes), int(rank), int(world), int(chunks)
        if balance == "edges":
            self.per_raw, self.bounds = block_bounds_by_edges(torch.bincount(dst, minlength=num_nodes), world)
        else:
            self.per_raw, self.bounds = block_bounds(num_nodes, world)
        self._bounds_t = torch.tensor(self.bounds, dtype=torch.int64)
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
        b = self._bounds_t.to(g.device)
        owner = torch.searchsorted(b[1:], g, right=True).clamp(max=self.world - 1)
        i = g - b[owner]
        return (i // self.cr) * (self.world * self.cr) + owner * self.cr + (i % self.cr)

    def local_rows(self, c: int):
        return c * self.cr, (c + 1) * self.cr

    def slab_rows(self, c: int):
        return c * self.world * self.cr, (c + 1) * self.world * self.cr

    @property
    def num_local_edges(self):
        return int(self.edge_ids.numel())


class FeaturePartitionPlan:
    """Index arithmetic of the feature partition: rank p owns rows [r0, r1) for the dense products and columns
    [p F/P, (p+1) F/P) of every aggregated matrix.  The whole graph is kept by every rank with node ids in the gathered
    numbering (owner * per + row_in_block), so the all-to-all's receive buffer IS the all-rows matrix."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int, rank: int, world: int, balance: str = "rows"):
        self.num_nodes, self.rank, self.world, self.chunks = int(num_nodes), int(rank), int(world), 1
        if balance == "edges":
            self.per_raw, self.bounds = block_bounds_by_edges(torch.bincount(dst, minlength=num_nodes), world)
        else:
            self.per_raw, self.bounds = block_bounds(num_nodes, world)
        self._bounds_t = torch.tensor(self.bounds, dtype=torch.int64)
        self.per = self.cr = self.per_raw
        self.r0, self.r1 = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.r1 - self.r0
        self.n_padded = self.per * world
        self.src_gathered = self.gathered_id(src)
        self.dst_gathered = self.gathered_id(dst)
        in_deg = torch.bincount(self.dst_gathered, minlength=self.n_padded).to(torch.float32)
        self.scale_full = 1.0 / in_deg.clamp(min=1.0)
        self.scale_local = self.scale_full[rank * self.per:(rank + 1) * self.per]

    def gathered_id(self, g: torch.Tensor) -> torch.Tensor:
        b = self._bounds_t.to(g.device)
        owner = torch.searchsorted(b[1:], g, right=True).clamp(max=self.world - 1)
        return owner * self.per + (g - b[owner])

    def col_range(self, feat: int):
        fc = feat // self.world
        return self.rank * fc, (self.rank + 1) * fc

    @property
    def num_local_edges(self):
        return int(self.src_gathered.numel())


class DistGCN(torch.nn.Module):
    """L-layer weighted-sum GCN over a partition.  Parameters are replicated (same seed on every rank)."""

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

    def grad_order(self):
        return [p for pair in zip(self.weights, self.biases) for p in pair]


class DistSAGEPool(torch.nn.Module):
    """L SAGEConv('pool') layers over a partition (the reference's layer, code/model.py:13-15, on the scaled graph):
    rst = h Ws^T + max_{u in in(v)} relu(h Wp^T + bp)[u] Wn^T + b, leaky_relu between layers, none after the last.
    Parameters replicated; per layer in the order fc_pool.weight, fc_pool.bias, fc_self.weight, fc_neigh.weight, bias."""

    def __init__(self, dims, seed: int = 0):
        super().__init__()
        gen = torch.Generator().manual_seed(seed)

        def uni(o, i):
            return torch.nn.Parameter((torch.rand(o, i, generator=gen) * 2 - 1) / math.sqrt(i))

        self.w_pool, self.b_pool, self.w_self, self.w_neigh, self.bias = (torch.nn.ParameterList() for _ in range(5))
        for i in range(len(dims) - 1):
            f, o = dims[i], dims[i + 1]
            self.w_pool.append(uni(f, f))
            self.b_pool.append(torch.nn.Parameter((torch.rand(f, generator=gen) * 2 - 1) / math.sqrt(f)))
            self.w_self.append(uni(o, f))
            self.w_neigh.append(uni(o, f))
            self.bias.append(torch.nn.Parameter(torch.zeros(o)))
        self.dims = list(dims)

    def grad_order(self):
        out = []
        for i in range(len(self.dims) - 1):
            out += [self.w_pool[i], self.b_pool[i], self.w_self[i], self.w_neigh[i], self.bias[i]]
        return out


class PartitionedGraph:
    """Device structures of one rank.  Row partition: in-edge CSR of the owned rows (gathered source ids) and its
    transpose.  Feature partition: in-edge and out-edge CSR of the whole graph in the gathered numbering; when the column
    slice of a rank is narrow (feat / world <= 64) each direction is additionally cut by SOURCE range into slabs whose slice of
    the feature matrix (rows x feat / world floats) is at most `slab_mb` MB, so that one aggregation pass gathers from an
    L2-resident slab (plagnn_spmm_*_slab).  Edge weights are stored once per structure in its CSR order (no edge-id
    indirection in the kernel)."""

    def __init__(self, plan, weight_global: torch.Tensor | None, build_csr, device, transposed: bool = True, feat: int | None = None,
                 slab_mb: float | None = None):
        import os
        self.plan = plan
        self.mode = "cols" if isinstance(plan, FeaturePartitionPlan) else "rows"
        if self.mode == "rows":
            s = plan.src_gathered.to(device=device, dtype=torch.int32)
            d = plan.dst_local.to(device=device, dtype=torch.int32)
            n_dst, w = plan.per, None if weight_global is None else weight_global[plan.edge_ids.to(weight_global.device)].to(device)
            self.scale = plan.scale_local.to(device)
        else:
            s = plan.src_gathered.to(device=device, dtype=torch.int32)
            d = plan.dst_gathered.to(device=device, dtype=torch.int32)
            n_dst, w = plan.n_padded, None if weight_global is None else weight_global.to(device)
            self.scale = plan.scale_local.to(device)
            self.scale_full = plan.scale_full.to(device)
        self.csc_slabs = self.w_csc_slabs = self.csr_t_slabs = self.w_csr_t_slabs = None
        n_slabs = 1
        if self.mode == "cols" and feat is not None and feat // plan.world <= 64:
            # off by default: measured on the 1 M / 100 M graph at 32 columns, 1 / 2 / 4 / 8 slabs -> 2.56 / 2.85 / 3.72 / 5.92 ms per
            # aggregation (the items get shorter and the kernel is latency-bound per item, not L2-miss-bound)
            slab_mb = float(os.environ.get("PLAGNN_DIST_SLAB_MB", "0")) if slab_mb is None else slab_mb
            if slab_mb > 0:
                n_slabs = max(1, math.ceil(plan.n_padded * (feat // plan.world) * 4 / (slab_mb * 2 ** 20)))
        self.n_slabs = n_slabs

        def slabs_of(key, other, n_key, n_other):
            """CSR structures with rows = key, entries = other restricted to `n_slabs` consecutive ranges of `other`."""
            rows_per = (n_other + n_slabs - 1) // n_slabs
            slab_id = other // rows_per
            out, ws = [], []
            for i in range(n_slabs):
                m = slab_id == i
                c = build_csr(key[m], other[m], n_key, False, num_other=n_other)
                out.append(c)
                ws.append(None if w is None else w[m][c.eids.long()].contiguous())
            return out, (None if w is None else ws)

        if n_slabs > 1:
            self.csc_slabs, self.w_csc_slabs = slabs_of(d, s, n_dst, plan.n_padded)
            if transposed:
                self.csr_t_slabs, self.w_csr_t_slabs = slabs_of(s, d, plan.n_padded, n_dst)
            self.csc = self.csc_slabs[0]                       # (identity of the direction for the backend's dispatch)
            self.csr_t = self.csr_t_slabs[0] if transposed else None
            self.w_csc = self.w_csr_t = None
            self.edge_weight = w
            return
        # rows = destination rows, entries = rows of the gathered matrix
        self.csc = build_csr(d, s, n_dst, False, num_other=plan.n_padded)
        self.w_csc = None if w is None else w[self.csc.eids.long()].contiguous()
        # transpose: rows = gathered (source) rows, entries = destination rows (only the sum reducer's backward walks it)
        self.csr_t = self.w_csr_t = None
        if transposed:
            self.csr_t = build_csr(s, d, plan.n_padded, False, num_other=n_dst)
            self.w_csr_t = None if w is None else w[self.csr_t.eids.long()].contiguous()
        self.edge_weight = w


def _mask_padded_rows(plan, d_out):
    """Rows past n_local are padding: they carry act(bias) in the forward pass and must not reach any gradient."""
    if plan.n_local < d_out.shape[0]:
        d_out[plan.n_local:] = 0
    return d_out


# ======================================================================================================================
# row partition
# ======================================================================================================================
def dist_gcn_forward_backward(model: DistGCN, pg, h0_local, backend, group=None, loss_grad_fn=None, act_leaky=True):
    """One forward + backward of the row-partitioned weighted-sum GCN.  h0_local: [per, F0] (rows past n_local are zero).
    Returns (out_local, grads) with grads = [W0, b0, W1, b1, ...] already summed over ranks."""
    if getattr(pg, "mode", "rows") == "cols":
        if len(model.weights) >= 2 and os.environ.get("PLAGNN_DIST_COLS_ALT", "1") != "0":
            return _gcn_cols_alt_forward_backward(model, pg, h0_local, backend, group, loss_grad_fn, act_leaky)
        return _gcn_cols_forward_backward(model, pg, h0_local, backend, group, loss_grad_fn, act_leaky)
    plan = pg.plan
    world, chunks = plan.world, plan.chunks
    n_layers = len(model.weights)
    dims = model.dims
    w_csc, w_csr_t = getattr(pg, "w_csc", pg.edge_weight), getattr(pg, "w_csr_t", pg.edge_weight)
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
                backend.spmm_rows(pg.csc, ("csc", c), a, b, t_full_prev, out_prev, w_csc, pg.scale,
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
    backend.spmm_rows(pg.csc, ("csc", -1), 0, plan.per, t_full_prev, out_last, w_csc, pg.scale,
                      model.biases[-1].detach(), act=False)
    outs.append(out_last)
    # ---------------- backward ----------------
    d_out = loss_grad_fn(out_last) if loss_grad_fn is not None else torch.ones_like(out_last)
    d_out = _mask_padded_rows(plan, d_out)
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
            backend.spmm_rows(pg.csr_t, ("csr_t", c), sa, sb, dzs, dt_partial, w_csr_t, None, None, act=False)
            a, b = plan.local_rows(c)
            pending.append(backend.reduce_scatter_chunk(backend.rows(dt_local, a, b), backend.rows(dt_partial, sa, sb),
                                                        world, group))
        backend.wait(pending)
        grads[2 * li] = backend.gemm_tn(dt_local, hs[li])                                  # [O, F]
        if li > 0:
            dz = backend.gemm_nn(dt_local, w)                                              # [per, F]
    return out_last, backend.all_reduce_grads(grads, world, group)


def dist_pool_forward_backward(model: DistSAGEPool, pg, h0_local, backend, group=None, loss_grad_fn=None):
    """One forward + backward of the partitioned SAGEConv-pool stack (either partition).  Returns (out_local, grads) with
    grads in DistSAGEPool.grad_order(), summed over ranks.

    Row partition: all-gather of m = relu(h Wp^T + bp) (chunk by chunk behind the projection), local max with arg over the
    gathered rows; backward: arg-scatter into a full-height partial dm, reduce-scatter slab by slab while the two
    weight-gradient products that do not need dm run."""
    cols = getattr(pg, "mode", "rows") == "cols"
    plan = pg.plan
    world, chunks = plan.world, plan.chunks
    n_layers = len(model.w_pool)
    dims = model.dims
    saved = []
    h = h0_local
    # ---------------- forward ----------------
    for li in range(n_layers):
        last = li + 1 == n_layers
        wp, bp, ws, wn, b = (model.w_pool[li].detach(), model.b_pool[li].detach(), model.w_self[li].detach(),
                             model.w_neigh[li].detach(), model.bias[li].detach())
        f = dims[li]
        if cols:
            m_local = backend.alloc(plan.per, f)
            backend.gemm_nt_bias_relu_into(h, wp, bp, m_local)
            m_col = backend.to_cols(m_local, world, group)
            fused = getattr(backend, "spmm_max_to_rows", None)
            if fused is not None:
                neigh, neigh_col, arg = fused(pg.csc, m_col, f, world, group)
            else:
                neigh_col, arg = backend.spmm_max(pg.csc, m_col)
                neigh = backend.to_rows(neigh_col, f, world, group)
            ctx = (neigh_col, arg)
        else:
            m_local = backend.alloc(plan.per, f)
            m_full = backend.alloc(plan.n_padded, f)
            pending = []
            for c in range(chunks):
                a, e = plan.local_rows(c)
                backend.gemm_nt_bias_relu_into(backend.rows(h, a, e), wp, bp, backend.rows(m_local, a, e))
                sa, sb = plan.slab_rows(c)
                pending.append(backend.all_gather_chunk(backend.rows(m_full, sa, sb), backend.rows(m_local, a, e), world, group))
            backend.wait(pending)
            neigh, arg = backend.spmm_max(pg.csc, m_full)
            ctx = (neigh, arg)
        out = backend.gemm2_nt(h, ws, neigh, wn, b, leaky=not last)
        saved.append((h, neigh, ctx, out))
        h = out
    out_last = h
    # ---------------- backward ----------------
    d_out = loss_grad_fn(out_last) if loss_grad_fn is not None else torch.ones_like(out_last)
    dz = _mask_padded_rows(plan, d_out)
    grads = [None] * (5 * n_layers)
    for li in reversed(range(n_layers)):
        last = li + 1 == n_layers
        wp, ws, wn = model.w_pool[li].detach(), model.w_self[li].detach(), model.w_neigh[li].detach()
        h, neigh, (neigh_agg, arg), out = saved[li]
        f = dims[li]
        drst = backend.act_backward(dz, None if last else out, None)
        dneigh = backend.gemm_nn(drst, wn)                                       # [per, F]
        if cols:
            dneigh_col = backend.to_cols(dneigh, world, group)
            dm_col = backend.max_scatter(dneigh_col, arg, neigh_agg, plan.n_padded)   # complete: every row of my columns
            grads[5 * li + 4] = backend.colsum(drst)
            grads[5 * li + 2] = backend.gemm_tn(drst, h)
            grads[5 * li + 3] = backend.gemm_tn(drst, neigh)
            dm_local = backend.to_rows(dm_col, f, world, group)
        else:
            dm_partial = backend.max_scatter(dneigh, arg, neigh_agg, plan.n_padded)   # relu' folded in: pooled value > 0
            dm_local = backend.alloc(plan.per, f)
            pending = []
            for c in range(chunks):
                sa, sb = plan.slab_rows(c)
                a, e = plan.local_rows(c)
                pending.append(backend.reduce_scatter_chunk(backend.rows(dm_local, a, e), backend.rows(dm_partial, sa, sb),
                                                            world, group))
            # these do not need dm: they run while the reduce-scatter is in flight
            grads[5 * li + 4] = backend.colsum(drst)
            grads[5 * li + 2] = backend.gemm_tn(drst, h)
            grads[5 * li + 3] = backend.gemm_tn(drst, neigh)
            backend.wait(pending)
        grads[5 * li + 0] = backend.gemm_tn(dm_local, h)
        grads[5 * li + 1] = backend.colsum(dm_local)
        if li > 0:
            dz = backend.gemm_nn2(drst, ws, dm_local, wp)                        # d h = drst Ws + dm Wp
    return out_last, backend.all_reduce_grads(grads, world, group)


# ======================================================================================================================
# feature partition, weighted sum
# ======================================================================================================================
def _spmm_cols_to_rows(backend, csx, key, x_col, w, scale, bias, act, feat, world, group):
    """Aggregate all rows of my columns, hand every owner its rows: fused block by block when the backend can (CUDA backend
    with the peer-memory exchange), else aggregation followed by the exchange."""
    fused = getattr(backend, "spmm_cols_to_rows", None)
    if fused is not None:
        return fused(csx, key, x_col, w, scale, bias, act, feat, world, group)
    return backend.to_rows(backend.spmm_cols(csx, key, x_col, w, scale, bias, act), feat, world, group)


def _gcn_cols_forward_backward(model: DistGCN, pg, h0_local, backend, group, loss_grad_fn, act_leaky):
    plan = pg.plan
    world = plan.world
    n_layers = len(model.weights)
    dims = model.dims
    hs, outs = [h0_local], []
    h = h0_local
    for li in range(n_layers):
        last = li + 1 == n_layers
        o = dims[li + 1]
        t_local = backend.alloc(plan.per, o)
        backend.gemm_nt_into(h, model.weights[li].detach(), t_local)
        t_col = backend.to_cols(t_local, world, group)
        c0, c1 = plan.col_range(o)
        h = _spmm_cols_to_rows(backend, pg.csc, "csc", t_col, pg.w_csc, pg.scale_full, model.biases[li].detach()[c0:c1],
                               act_leaky and not last, o, world, group)
        outs.append(h)
        if not last:
            hs.append(h)
    out_last = h
    d_out = loss_grad_fn(out_last) if loss_grad_fn is not None else torch.ones_like(out_last)
    dz = _mask_padded_rows(plan, d_out)
    grads = [None] * (2 * n_layers)
    for li in reversed(range(n_layers)):
        last = li + 1 == n_layers
        o = dims[li + 1]
        dzb = backend.act_backward(dz, outs[li] if (not last and act_leaky) else None, None)
        grads[2 * li + 1] = backend.colsum(dzb)
        dzs = backend.act_backward(dzb, None, pg.scale)
        dzs_col = backend.to_cols(dzs, world, group)
        dt_local = _spmm_cols_to_rows(backend, pg.csr_t, "csr_t", dzs_col, pg.w_csr_t, None, None, False, o, world, group)
        grads[2 * li] = backend.gemm_tn(dt_local, hs[li])
        if li > 0:
            dz = backend.gemm_nn(dt_local, model.weights[li].detach())
    return out_last, backend.all_reduce_grads(grads, world, group)


def _gcn_cols_alt_forward_backward(model: DistGCN, pg, h0_local, backend, group, loss_grad_fn, act_leaky):
    """Feature partition with the layers taken in pairs: layer 2j transforms first (X W, then aggregate), layer 2j + 1 aggregates
    first (aggregate, then multiply by W) — the order DGL's GraphConv itself picks when in_feats <= out_feats.  The two
    aggregations of a pair then follow each other in the column layout (bias and activation are per column / elementwise and
    need no exchange), so a pair costs two exchanges per direction instead of four: rows -> columns before the first
    aggregation, columns -> rows after the second.  Same function as `_gcn_cols_forward_backward` up to fp32 rounding
    ((A h) W against A (h W))."""
    plan = pg.plan
    world = plan.world
    n_layers = len(model.weights)
    dims = model.dims
    W = [w.detach() for w in model.weights]
    B = [b.detach() for b in model.biases]
    saved = [None] * n_layers
    h = h0_local
    li = 0
    while li < n_layers:
        last = li + 1 == n_layers
        o = dims[li + 1]
        c0, c1 = plan.col_range(o)
        t_local = backend.alloc(plan.per, o)
        backend.gemm_nt_into(h, W[li], t_local)
        t_col = backend.to_cols(t_local, world, group)
        if last:                                   # an odd layer count: the last layer stands alone
            out = _spmm_cols_to_rows(backend, pg.csc, "csc", t_col, pg.w_csc, pg.scale_full, B[li][c0:c1], False, o, world, group)
            saved[li] = ("single", h, None)
            h = out
            li += 1
            continue
        z_col = backend.spmm_cols(pg.csc, "csc", t_col, pg.w_csc, pg.scale_full, B[li][c0:c1], act_leaky)      # all rows x my columns
        saved[li] = ("first", h, z_col)
        last2 = li + 2 == n_layers
        # (aggregation + exchange fused block by block when the backend can: PLAGNN_DIST_OVERLAP)
        a_rows = _spmm_cols_to_rows(backend, pg.csc, "csc", z_col, pg.w_csc, pg.scale_full, None, False, o, world, group)
        out = backend.gemm_nt_bias_act(a_rows, W[li + 1], B[li + 1], act_leaky and not last2)
        saved[li + 1] = ("second", a_rows, out)
        h = out
        li += 2
    out_last = h
    d_out = loss_grad_fn(out_last) if loss_grad_fn is not None else torch.ones_like(out_last)
    d = _mask_padded_rows(plan, d_out)
    grads = [None] * (2 * n_layers)
    li = n_layers - 1
    while li >= 0:
        kind, x_in, y = saved[li]
        last = li + 1 == n_layers
        if kind == "single":
            grads[2 * li + 1] = backend.colsum(d)
            dzs_col = backend.to_cols(backend.act_backward(d, None, pg.scale), world, group)
            dt_local = _spmm_cols_to_rows(backend, pg.csr_t, "csr_t", dzs_col, pg.w_csr_t, None, None, False, dims[li + 1], world, group)
            grads[2 * li] = backend.gemm_tn(dt_local, x_in)
            if li > 0:
                d = backend.gemm_nn(dt_local, W[li])
            li -= 1
            continue
        # the pair (li - 1, li): out = act(a_rows W_li^T + b_li), a = A z, z = act(A (h W_{li-1}^T) + b_{li-1})
        dzb = backend.act_backward(d, y if (not last and act_leaky) else None, None)
        grads[2 * li + 1] = backend.colsum(dzb)
        grads[2 * li] = backend.gemm_tn(dzb, x_in)
        da = backend.gemm_nn(dzb, W[li])                                       # my rows x dims[li]
        das_col = backend.to_cols(backend.act_backward(da, None, pg.scale), world, group)
        dz_col = backend.spmm_cols(pg.csr_t, "csr_t", das_col, pg.w_csr_t, None, None, False)      # all rows x my columns
        lp = li - 1
        _, h_in, z_col = saved[lp]
        o = dims[lp + 1]
        c0, c1 = plan.col_range(o)
        dzb_col = backend.act_backward(dz_col, z_col if act_leaky else None, None)
        gb = torch.zeros(o, device=dzb_col.device, dtype=dzb_col.dtype)
        gb[c0:c1] = backend.colsum(dzb_col)                                     # complete for my columns, zero elsewhere: the sum
        grads[2 * lp + 1] = gb                                                  # over ranks assembles the vector
        dzs_col = backend.act_backward(dzb_col, None, pg.scale_full)
        dt_local = _spmm_cols_to_rows(backend, pg.csr_t, "csr_t", dzs_col, pg.w_csr_t, None, None, False, o, world, group)
        grads[2 * lp] = backend.gemm_tn(dt_local, h_in)
        if lp > 0:
            d = backend.gemm_nn(dt_local, W[lp])
        li -= 2
    return out_last, backend.all_reduce_grads(grads, world, group)


# ======================================================================================================================
# CUDA backend: the library's kernels + its NCCL wrappers
# ======================================================================================================================
class NcclComm:
    """An NCCL communicator of the library's own (plagnn_nccl_comm_init).  The 128-byte id travels from rank 0 through
    the already initialised torch.distributed group (any backend); after that the data path does not touch
    torch.distributed.  max_ctas > 0 caps the CTAs NCCL may use (the exchange runs beside the aggregation)."""

    def __init__(self, rank: int, world: int, device, max_ctas: int = 0):
        from . import _lib
        lib = _lib.load()
        self.rank, self.world = rank, world
        idbuf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.plagnn_nccl_get_unique_id(idbuf), "nccl_get_unique_id")
        box = [bytes(idbuf)]
        dist.broadcast_object_list(box, src=0)
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
        comm = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.plagnn_nccl_comm_init(idbuf, rank, world, int(max_ctas), ctypes.byref(comm)), "nccl_comm_init")
        self.handle = comm
        self._lib = lib

    def destroy(self):
        if self.handle:
            self._lib.plagnn_nccl_comm_destroy(self.handle)
            self.handle = None


class _WindowView:
    """CUDA array interface over a slice of a peer-exchange window, so that torch can view library-owned device memory."""

    def __init__(self, ptr: int, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self._owner = owner


class P2PExchange:
    """Peer-memory exchange of the feature partition (plagnn_p2p_*): one window per rank, opened by every peer over CUDA IPC;
    the all-to-all steps become one kernel that stores straight into the owners' windows plus a flag wait.  `region_bytes` =
    rows_per_rank * feat * 4 (the size of one exchanged matrix).  The window is a ring of `regions` such matrices used in turn:
    a result may be kept (saved activation) for up to regions - 1 further exchanges without a copy — a step of the 2-layer
    models makes 8 — and a peer can only reach the exchange that overwrites a region after it has seen this rank's flags for all
    exchanges in between, which this rank publishes in program order after its reads."""

    def __init__(self, region_bytes: int, rank: int, world: int, device, regions: int = 16):
        from . import _lib
        self._lib, self.lib = _lib, _lib.load()
        self.rank, self.world, self.device = rank, world, device
        self.region_bytes = (int(region_bytes) + 255) // 256 * 256
        self.regions = int(regions)
        self.handle = ctypes.c_void_p()
        hbuf = (ctypes.c_ubyte * 64)()
        # every step is agreed on by all ranks (a rank that failed alone would leave the others waiting in a collective)
        err = None
        try:
            with torch.cuda.device(device):
                _lib.check(self.lib.plagnn_p2p_create(self.regions * self.region_bytes, rank, world, hbuf, ctypes.byref(self.handle)),
                           "p2p_create")
        except Exception as ex:
            err = repr(ex)
        got = [None] * world
        dist.all_gather_object(got, (err, bytes(hbuf)))
        if any(e is not None for e, _ in got):
            self._free()
            raise _lib.PlagnnError("peer-memory window: " + "; ".join(str(e) for e, _ in got if e is not None))
        allh = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(h for _, h in got))
        try:
            with torch.cuda.device(device):
                _lib.check(self.lib.plagnn_p2p_attach(self.handle, allh), "p2p_attach")
        except Exception as ex:
            err = repr(ex)
        got = [None] * world
        dist.all_gather_object(got, err)            # also: every window is open everywhere before the first store
        if any(e is not None for e in got):
            self._free()
            raise _lib.PlagnnError("peer-memory window: " + "; ".join(str(e) for e in got if e is not None))
        self.base = int(self.lib.plagnn_p2p_window(self.handle))
        self.seq = 0

    def _free(self):
        if self.handle:
            self.lib.plagnn_p2p_destroy(self.handle)
            self.handle = None

    def _region(self, seq):
        return (seq % self.regions) * self.region_bytes

    def exchange(self, src: torch.Tensor, rows: int, feat: int, mode: int, stream) -> torch.Tensor:
        """Sends `src` (mode 0: [rows x feat] -> all rows x my columns; mode 1: [world*rows x feat/world] -> my rows x all
        columns), waits for the peers' parts and returns a view of the window region that now holds the result."""
        self.seq += 1
        off = self._region(self.seq)
        self._lib.check(self.lib.plagnn_p2p_send(self.handle, src.data_ptr(), src.stride(0), rows, feat, mode, off, self.seq, stream),
                        "p2p_send")
        self._lib.check(self.lib.plagnn_p2p_wait(self.handle, self.seq, stream), "p2p_wait")
        shape = (self.world * rows, feat // self.world) if mode == 0 else (rows, feat)
        return torch.as_tensor(_WindowView(self.base + off, shape, self), device=self.device)

    def begin(self):
        """A new exchange made of several parts: returns (seq, window offset)."""
        self.seq += 1
        return self.seq, self._region(self.seq)

    def send_part(self, src, rows, feat, mode, row_begin, row_end, publish, seq, off, stream):
        self._lib.check(self.lib.plagnn_p2p_send_part(self.handle, src.data_ptr(), src.stride(0), rows, feat, mode, row_begin, row_end,
                                                      publish, off, seq, stream), "p2p_send_part")

    def finish(self, seq, off, rows, feat, mode, stream) -> torch.Tensor:
        self._lib.check(self.lib.plagnn_p2p_wait(self.handle, seq, stream), "p2p_wait")
        shape = (self.world * rows, feat // self.world) if mode == 0 else (rows, feat)
        return torch.as_tensor(_WindowView(self.base + off, shape, self), device=self.device)

    def error(self) -> int:
        return int(self.lib.plagnn_p2p_error(self.handle))

    def destroy(self):
        if self.handle:
            torch.cuda.synchronize()
            dist.barrier()             # nobody is still storing into a window that is about to go away
            self._free()


class CudaBackend:
    """Compute + collectives on the CUDA kernels / NCCL wrappers.  (Tests use a CPU twin with the same method names.)
    comm: NcclComm or None (world 1).  skip_comm = True leaves every collective out (timing of the compute alone: the
    difference to the full step is the exposed exchange time; results are then meaningless)."""

    def __init__(self, pg: PartitionedGraph, comm: NcclComm | None = None, p2p: "P2PExchange | None" = None):
        from . import _lib, ops
        self.ops, self._lib, self.lib = ops, _lib, _lib.load()
        self.comm = comm
        self.p2p = p2p                 # feature partition: peer-memory exchange instead of pack + NCCL all-to-all + unpack
        self.pg = pg
        self.skip_comm = False
        self.overlap = False           # feature partition: aggregation and exchange fused block by block (set below)
        # highest priority: when an SM slot frees up, the exchange kernel's CTAs are placed before the pending CTAs of the
        # aggregation launched ahead of it — otherwise the collective only starts once that whole grid has been scheduled
        self.comm_stream = torch.cuda.Stream(priority=-1) if (comm is not None or p2p is not None) else None
        plan = pg.plan
        self.ranges = {}
        if pg.mode == "rows":
            for c in range(plan.chunks):
                self.ranges[("csc", c)] = ops.plan_range(pg.csc, *plan.local_rows(c))
                if pg.csr_t is not None:
                    self.ranges[("csr_t", c)] = ops.plan_range(pg.csr_t, *plan.slab_rows(c))
            self.ranges[("csc", -1)] = ops.plan_range(pg.csc, 0, plan.per)
        elif p2p is not None and plan.world > 1 and pg.csc_slabs is None:
            # feature partition: the aggregation is launched per owner's row block, so that each block can leave for its owner
            # while the next one is aggregated
            for q in range(plan.world):
                self.ranges[("csc", q)] = ops.plan_range(pg.csc, q * plan.per, (q + 1) * plan.per)
                if pg.csr_t is not None:
                    self.ranges[("csr_t", q)] = ops.plan_range(pg.csr_t, q * plan.per, (q + 1) * plan.per)
            # off by default: measured at 2 GPUs / 32 columns per rank, the per-block launches cost the aggregation more
            # (10.8 vs 10.2 ms) than the hidden stores save: 13.0 vs 12.6 ms per step (weighted sum), 10.7 vs 10.5 (max)
            self.overlap = os.environ.get("PLAGNN_DIST_OVERLAP", "0") == "1"

    def alloc(self, rows, cols):
        return self.ops.alloc(rows, cols, torch.cuda.current_device())

    @staticmethod
    def rows(t, a, b):
        return t[a:b]

    # ---- dense ----------------------------------------------------------------------------------------------------
    def gemm_nt_into(self, a, w, out):
        ops = self.ops
        a = ops.aligned(a)
        ops.gemm(a.shape[0], w.shape[0], [(a, 0, ops.aligned(w), 0, a.shape[1])], out=out)

    def gemm_nt_bias_relu_into(self, a, w, bias, out):
        ops = self.ops
        a = ops.aligned(a)
        ops.gemm(a.shape[0], w.shape[0], [(a, 0, ops.aligned(w), 0, a.shape[1])], bias=bias, act=ops.ACT_RELU, out=out)

    def gemm_nt_bias_act(self, a, w, bias, leaky):
        ops = self.ops
        a = ops.aligned(a)
        return ops.gemm(a.shape[0], w.shape[0], [(a, 0, ops.aligned(w), 0, a.shape[1])], bias=bias,
                        act=ops.ACT_LEAKY if leaky else ops.ACT_NONE)

    def gemm2_nt(self, h, ws, neigh, wn, bias, leaky):
        ops = self.ops
        h, neigh = ops.aligned(h), ops.aligned(neigh)
        return ops.gemm(h.shape[0], ws.shape[0], [(h, 0, ops.aligned(ws), 0, h.shape[1]), (neigh, 0, ops.aligned(wn), 0, h.shape[1])],
                        bias=bias, act=ops.ACT_LEAKY if leaky else ops.ACT_NONE)

    def gemm_nn(self, a, w):
        ops = self.ops
        a = ops.aligned(a)
        return ops.gemm(a.shape[0], w.shape[1], [(a, 0, ops.aligned(w), 1, a.shape[1])])

    def gemm_nn2(self, a0, w0, a1, w1):
        ops = self.ops
        a0, a1 = ops.aligned(a0), ops.aligned(a1)
        return ops.gemm(a0.shape[0], w0.shape[1], [(a0, 0, ops.aligned(w0), 1, a0.shape[1]), (a1, 0, ops.aligned(w1), 1, a1.shape[1])])

    def gemm_tn(self, a, b):
        ops = self.ops
        a, b = ops.aligned(a), ops.aligned(b)
        out = torch.empty((a.shape[1], b.shape[1]), device=a.device, dtype=torch.float32)
        return ops.gemm(a.shape[1], b.shape[1], [(a, 1, b, 1, a.shape[0])], out=out)

    def act_backward(self, dy, y, scale):
        ops = self.ops
        if y is None and scale is None:
            return dy
        return ops.act_backward(dy, y, ops.ACT_LEAKY if y is not None else ops.ACT_NONE, row_scale=scale)

    def colsum(self, x):
        return self.ops.colsum(self.ops.aligned(x))

    # ---- aggregation ----------------------------------------------------------------------------------------------
    def spmm_rows(self, csx, key, a, b, x, out, w, scale, bias, act):
        ops = self.ops
        ops.spmm_sum_rows(csx, self.ranges[key], x, out, w=w, scale=scale, bias=bias,
                          act=ops.ACT_LEAKY if act else ops.ACT_NONE, w_in_csr_order=True)

    def spmm_cols(self, csx, key, x_col, w, scale, bias, act):
        ops, pg = self.ops, self.pg
        a = ops.ACT_LEAKY if act else ops.ACT_NONE
        slabs = pg.csc_slabs if key == "csc" else pg.csr_t_slabs
        if slabs is not None:                      # narrow column slice: one pass per L2-sized source slab
            return ops.spmm_sum_slabs(slabs, x_col, ws=pg.w_csc_slabs if key == "csc" else pg.w_csr_t_slabs, scale=scale,
                                      bias=bias, act=a)
        return ops.spmm_sum(csx, x_col, w=w, scale=scale, bias=bias, act=a, w_in_csr_order=True)

    def spmm_max(self, csc, x):
        if self.pg.csc_slabs is not None:
            return self.ops.spmm_max_slabs(self.pg.csc_slabs, x)
        return self.ops.spmm_max_fwd(csc, x)

    def _blocks_to_rows(self, launch_block, out_col, feat, world):
        """Feature partition, aggregation + exchange fused block by block: `launch_block(q)` aggregates the rows owned by rank
        q into `out_col`; as soon as a block is done its rows leave for their owner on the exchange stream (peer-memory
        store, flag for that owner only) while the next block is aggregated.  Returns my rows x all columns."""
        ops, p2p = self.ops, self.p2p
        rows = out_col.shape[0] // world
        cur = torch.cuda.current_stream()
        seq, off = p2p.begin()
        cst = ctypes.c_void_p(self.comm_stream.cuda_stream)
        last = None
        for q in range(world):
            launch_block(q)
            done = torch.cuda.Event()
            done.record(cur)
            self.comm_stream.wait_event(done)
            p2p.send_part(out_col, rows, feat, 1, q * rows, (q + 1) * rows, 2, seq, off, cst)
            last = torch.cuda.Event()
            last.record(self.comm_stream)
        res = p2p.finish(seq, off, rows, feat, 1, ops._stream())
        cur.wait_event(last)          # my own stores are finished before out_col can be reused (and before the next send kernel)
        return res

    def spmm_cols_to_rows(self, csx, key, x_col, w, scale, bias, act, feat, world, group):
        ops, pg = self.ops, self.pg
        if not (self.overlap and self.p2p is not None and world > 1 and not self.skip_comm):
            return self.to_rows(self.spmm_cols(csx, key, x_col, w, scale, bias, act), feat, world, group)
        a = ops.ACT_LEAKY if act else ops.ACT_NONE
        x_col = ops.aligned(x_col)
        out_col = ops.alloc(csx.num_rows, x_col.shape[1], x_col.device)
        return self._blocks_to_rows(lambda q: ops.spmm_sum_rows(csx, self.ranges[(key, q)], x_col, out_col, w=w, scale=scale, bias=bias,
                                                                act=a, w_in_csr_order=True), out_col, feat, world)

    def spmm_max_to_rows(self, csc, x_col, feat, world, group):
        """Returns (neigh my rows x all columns, neigh all rows x my columns, arg all rows x my columns)."""
        ops = self.ops
        if not (self.overlap and self.p2p is not None and world > 1 and not self.skip_comm):
            neigh_col, arg = self.spmm_max(csc, x_col)
            return self.to_rows(neigh_col, feat, world, group), neigh_col, arg
        x_col = ops.aligned(x_col)
        f = x_col.shape[1]
        neigh_col = ops.alloc(csc.num_rows, f, x_col.device)
        arg = ops.alloc(csc.num_rows, f, x_col.device, dtype=torch.int32)
        neigh = self._blocks_to_rows(lambda q: ops.spmm_max_fwd_rows(csc, self.ranges[("csc", q)], x_col, neigh_col, arg), neigh_col, feat, world)
        return neigh, neigh_col, arg

    def max_scatter(self, dneigh, arg, neigh, n_src):
        return self.ops.spmm_max_bwd(dneigh, arg, neigh, n_src)

    # ---- exchange -------------------------------------------------------------------------------------------------
    @staticmethod
    def _pitch(t):
        return t.stride(0)

    def _on_comm_stream(self, fn):
        """Runs fn(stream_ptr) on the exchange stream after everything enqueued so far on the compute stream; returns the
        event the compute stream must wait for before it touches the result."""
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.comm_stream.wait_event(ready)
        fn(ctypes.c_void_p(self.comm_stream.cuda_stream))
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        return done

    def all_gather_chunk(self, slab, local_chunk, world, group):
        if world == 1:
            slab.copy_(local_chunk)
            return None
        if self.skip_comm:
            return None
        assert self._pitch(slab) == self._pitch(local_chunk)
        rows, pitch = local_chunk.shape[0], self._pitch(local_chunk)
        return self._on_comm_stream(lambda st: self._lib.check(self.lib.plagnn_nccl_allgather_rows(
            local_chunk.data_ptr(), slab.data_ptr(), rows, pitch, self.comm.handle, st), "nccl_allgather_rows"))

    def reduce_scatter_chunk(self, out_chunk, slab, world, group):
        if world == 1:
            out_chunk.copy_(slab)
            return None
        if self.skip_comm:
            return None
        assert self._pitch(slab) == self._pitch(out_chunk)
        rows, pitch = out_chunk.shape[0], self._pitch(out_chunk)
        return self._on_comm_stream(lambda st: self._lib.check(self.lib.plagnn_nccl_reducescatter_rows(
            slab.data_ptr(), out_chunk.data_ptr(), rows, pitch, self.comm.handle, st), "nccl_reducescatter_rows"))

    @staticmethod
    def wait(pending):
        cur = torch.cuda.current_stream()
        for ev in pending:
            if ev is not None:
                cur.wait_event(ev)        # the compute stream waits for the exchange stream; the host does not block

    def to_cols(self, x_local, world, group):
        """[per x F] (my rows, all columns) -> [world * per x F / world] (all rows, my columns)."""
        ops = self.ops
        x_local = ops.aligned(x_local)
        rows, feat = x_local.shape
        fc = feat // world
        st = ops._stream()
        if self.p2p is not None and world > 1 and not self.skip_comm:
            return self.p2p.exchange(x_local, rows, feat, 0, st)       # a view of the window: consumed by the next kernel
        send = torch.empty((world * rows, fc), device=x_local.device, dtype=torch.float32)
        self._lib.check(self.lib.plagnn_cols_pack(x_local.data_ptr(), x_local.stride(0), rows, feat, world, send.data_ptr(), st),
                        "cols_pack")
        if world == 1 or self.skip_comm:
            return send
        recv = torch.empty_like(send)
        self._lib.check(self.lib.plagnn_nccl_alltoall_blocks(send.data_ptr(), recv.data_ptr(), rows * fc, world,
                                                             self.comm.handle, st), "nccl_alltoall_blocks")
        return recv

    def to_rows(self, x_col, feat, world, group):
        """[world * per x F / world] (all rows, my columns) -> [per x F] (my rows, all columns)."""
        ops = self.ops
        fc = feat // world
        rows = x_col.shape[0] // world
        st = ops._stream()
        if self.p2p is not None and world > 1 and not self.skip_comm:
            if x_col.stride(0) % 4 or x_col.data_ptr() % 16:
                x_col = x_col.contiguous()
            # callers keep this matrix (layer input / saved activation): it stays valid for regions - 1 further exchanges
            return self.p2p.exchange(x_col, rows, feat, 1, st)
        if x_col.stride(0) != fc:
            x_col = x_col.contiguous()
        recv = x_col
        if world > 1 and not self.skip_comm:
            recv = torch.empty((world * rows, fc), device=x_col.device, dtype=torch.float32)
            self._lib.check(self.lib.plagnn_nccl_alltoall_blocks(x_col.data_ptr(), recv.data_ptr(), rows * fc, world,
                                                                 self.comm.handle, st), "nccl_alltoall_blocks")
        out = ops.alloc(rows, feat, x_col.device)
        self._lib.check(self.lib.plagnn_cols_unpack(recv.data_ptr(), rows, feat, world, out.data_ptr(), out.stride(0), st),
                        "cols_unpack")
        return out

    def all_reduce_grads(self, grads, world, group):
        if world == 1 or self.skip_comm:
            return grads
        flat = torch.cat([g.reshape(-1) for g in grads])
        self._lib.check(self.lib.plagnn_nccl_allreduce(flat.data_ptr(), flat.numel(), self.comm.handle, self.ops._stream()),
                        "nccl_allreduce")
        out, off = [], 0
        for g in grads:
            out.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return out
