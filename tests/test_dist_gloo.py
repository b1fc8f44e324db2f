"""N > 1 host logic on CPU: world_size-2 and -3 gloo runs of the partitioned choreographies (row partition and feature
partition, weighted-sum and max-pool reducers) with a CPU compute backend (oracle kernels), compared with the
single-process oracle model.  No CUDA involved."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plagnn_oracle as orc
from plagnn_b200 import synth
from plagnn_b200.dist import (DistGCN, DistSAGEPool, FeaturePartitionPlan, RowPartitionPlan, block_bounds,
                              block_bounds_by_edges, dist_gcn_forward_backward, dist_pool_forward_backward)


class _Csx:
    def __init__(self, key, other, n_rows):
        self.indptr, self.indices, self.eids = orc.coo_to_csc(other.numpy(), key.numpy(), n_rows)
        self.num_rows = n_rows


class CpuBackend:
    """Same method names as plagnn_b200.dist.CudaBackend, computed with torch-CPU + the oracle's C SpMM (synchronous)."""

    def alloc(self, rows, cols): return torch.zeros(rows, cols)
    @staticmethod
    def rows(t, a, b): return t[a:b]
    def gemm_nt_into(self, a, w, out): out.copy_(a @ w.t())
    def gemm_nn(self, a, w): return a @ w
    def gemm_tn(self, a, b): return a.t() @ b
    def colsum(self, x): return x.sum(0)

    def gemm_nt_bias_relu_into(self, a, w, bias, out): out.copy_(torch.relu(a @ w.t() + bias))
    def gemm_nn2(self, a0, w0, a1, w1): return a0 @ w0 + a1 @ w1

    def gemm_nt_bias_act(self, a, w, bias, leaky):
        y = a @ w.t() + bias
        return torch.where(y > 0, y, 0.01 * y) if leaky else y

    def gemm2_nt(self, h, ws, neigh, wn, bias, leaky):
        r = h @ ws.t() + neigh @ wn.t() + bias
        return torch.nn.functional.leaky_relu(r) if leaky else r

    def spmm_max(self, csc, x):
        return orc.spmm_max_c(csc.indptr, csc.indices, x.contiguous())

    def max_scatter(self, dneigh, arg, neigh, n_src):
        return orc.spmm_max_bwd_c(arg, (dneigh * (neigh > 0)).contiguous(), n_src)

    def spmm_cols(self, csx, key, x_col, w, scale, bias, act):
        r = orc.spmm_sum_c(csx.indptr, csx.indices, x_col.contiguous(), eids=None, w=w, scale=scale)
        if bias is not None:
            r = r + bias
        return torch.nn.functional.leaky_relu(r) if act else r

    def to_cols(self, x_local, world, group):
        rows, feat = x_local.shape
        fc = feat // world
        send = x_local.reshape(rows, world, fc).permute(1, 0, 2).contiguous()      # [world][rows][fc]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        return recv.reshape(world * rows, fc)

    def to_rows(self, x_col, feat, world, group):
        fc = feat // world
        rows = x_col.shape[0] // world
        recv = torch.empty(world, rows, fc)
        dist.all_to_all_single(recv, x_col.contiguous().reshape(world, rows, fc), group=group)
        return recv.permute(1, 0, 2).reshape(rows, feat).contiguous()

    def spmm_rows(self, csx, key, a, b, x, out, w, scale, bias, act):
        r = orc.spmm_sum_c(csx.indptr[a:b + 1], csx.indices, x.contiguous(), eids=None, w=w,
                           scale=None if scale is None else scale[a:b].contiguous())
        if bias is not None:
            r = r + bias
        out[a:b] = torch.nn.functional.leaky_relu(r) if act else r

    def act_backward(self, dy, y, scale):
        if y is not None:
            dy = dy * torch.where(y > 0, 1.0, 0.01)
        if scale is not None:
            dy = dy * scale.unsqueeze(1)
        return dy

    def all_gather_chunk(self, slab, local_chunk, world, group):
        parts = [torch.empty_like(local_chunk) for _ in range(world)]
        dist.all_gather(parts, local_chunk.contiguous(), group=group)
        slab.copy_(torch.cat(parts))
        return None

    def reduce_scatter_chunk(self, out_chunk, slab, world, group):
        full = slab.clone()
        dist.all_reduce(full, group=group)
        n = out_chunk.shape[0]
        r = dist.get_rank(group)
        out_chunk.copy_(full[r * n:(r + 1) * n])
        return None

    @staticmethod
    def wait(pending): pass

    def all_reduce_grads(self, grads, world, group):
        out = []
        for g in grads:
            g = g.clone()
            dist.all_reduce(g, group=group)
            out.append(g)
        return out


class _PG:
    """CPU twin of plagnn_b200.dist.PartitionedGraph (weights stored per direction in CSR order)."""

    def __init__(self, plan, weight):
        self.plan = plan
        self.mode = "cols" if isinstance(plan, FeaturePartitionPlan) else "rows"
        if self.mode == "rows":
            d, n_dst, w = plan.dst_local, plan.per, weight[plan.edge_ids]
        else:
            d, n_dst, w = plan.dst_gathered, plan.n_padded, weight
            self.scale_full = plan.scale_full
        self.csc = _Csx(d, plan.src_gathered, n_dst)
        self.csr_t = _Csx(plan.src_gathered, d, plan.n_padded)
        self.w_csc = w[torch.as_tensor(self.csc.eids, dtype=torch.long)].contiguous()
        self.w_csr_t = w[torch.as_tensor(self.csr_t.eids, dtype=torch.long)].contiguous()
        self.edge_weight = w
        self.scale = plan.scale_local


def _problem(n=403, e=6000, f=None):
    f = DIMS[0] if f is None else f
    sg = synth.scaled_graph(n, e, seed=5, max_degree=200)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    return sg, x


# every width a multiple of 4 * world for world in (2, 3); PLAGNN_TEST_DIMS (inherited by the spawned ranks) overrides
DIMS = [int(v) for v in os.environ["PLAGNN_TEST_DIMS"].split(",")] if os.environ.get("PLAGNN_TEST_DIMS") else [24, 24, 12]


def _make_plan(sg, n, rank, world, mode, balance):
    if mode == "cols":
        return FeaturePartitionPlan(sg.src, sg.dst, n, rank, world, balance=balance)
    return RowPartitionPlan(sg.src, sg.dst, n, rank, world, chunks=3, balance=balance)


def _worker(rank, world, port, out_dir, mode, reducer, balance):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sg, x = _problem()
    n, f = sg.num_nodes, x.shape[1]
    plan = _make_plan(sg, n, rank, world, mode, balance)
    pg = _PG(plan, sg.weight)
    h0 = torch.zeros(plan.per, f)
    h0[:plan.n_local] = x[plan.r0:plan.r1]
    # no row mask here: padded rows (act(bias) in the forward pass) must be kept out of the gradients by the library
    if reducer == "sum":
        model = DistGCN(DIMS, seed=3)
        with torch.no_grad():
            for b in model.biases:
                b.add_(0.05)                      # non-zero biases: padded rows are not zero after the first layer
        out, grads = dist_gcn_forward_backward(model, pg, h0, CpuBackend(), None, lambda o: o / n)
    else:
        model = DistSAGEPool(DIMS, seed=3)
        with torch.no_grad():
            for b in model.bias:
                b.add_(0.05)
        out, grads = dist_pool_forward_backward(model, pg, h0, CpuBackend(), None, lambda o: o / n)
    torch.save({"out": out[:plan.n_local], "grads": grads, "r0": plan.r0, "r1": plan.r1},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


class _PoolStackRef(torch.nn.Module):
    """The oracle's SAGEConv-pool layers stacked like DistSAGEPool (leaky_relu between layers, none after the last)."""

    def __init__(self, model: DistSAGEPool):
        super().__init__()
        self.convs = torch.nn.ModuleList()
        for i in range(len(model.dims) - 1):
            c = orc.SAGEConvPoolRef(model.dims[i], model.dims[i + 1], "pool")
            with torch.no_grad():
                c.fc_pool.weight.copy_(model.w_pool[i]); c.fc_pool.bias.copy_(model.b_pool[i])
                c.fc_self.weight.copy_(model.w_self[i]); c.fc_neigh.weight.copy_(model.w_neigh[i]); c.bias.copy_(model.bias[i])
            self.convs.append(c)

    def forward(self, g, x):
        h = x
        for i, c in enumerate(self.convs):
            h = c(g, h)
            if i + 1 < len(self.convs):
                h = torch.nn.functional.leaky_relu(h)
        return h

    def grads(self):
        out = []
        for c in self.convs:
            out += [c.fc_pool.weight.grad, c.fc_pool.bias.grad, c.fc_self.weight.grad, c.fc_neigh.weight.grad, c.bias.grad]
        return out


def _reference(reducer="sum"):
    sg, x = _problem()
    n, f = sg.num_nodes, x.shape[1]
    g = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    if reducer == "sum":
        model = DistGCN(DIMS, seed=3)
        ref = orc.GCNSumRef(DIMS)
        with torch.no_grad():
            for lin, w, b in zip(ref.lins, model.weights, model.biases):
                lin.weight.copy_(w); lin.bias.copy_(b + 0.05)
        scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
        out = ref(g, x, sg.weight, scale)
        (0.5 * (out ** 2).sum() / n).backward()
        return out.detach(), [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]
    model = DistSAGEPool(DIMS, seed=3)
    with torch.no_grad():
        for b in model.bias:
            b.add_(0.05)
    ref = _PoolStackRef(model)
    out = ref(g, x)
    (0.5 * (out ** 2).sum() / n).backward()
    return out.detach(), ref.grads()


def test_block_bounds_and_plan_cover_every_edge_once():
    sg, _ = _problem()
    n = sg.num_nodes
    for world in (1, 2, 3, 8):
        per, b = block_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and all(b[i + 1] - b[i] <= per for i in range(world))
        seen = torch.zeros(sg.src.numel(), dtype=torch.int32)
        gathered = []
        for r in range(world):
            plan = RowPartitionPlan(sg.src, sg.dst, n, r, world, chunks=3)
            seen[plan.edge_ids] += 1
            # chunk-major gathered numbering: the all-gather of local chunk c fills slab c
            own = torch.arange(plan.r0, plan.r1)
            gid = plan.gathered_id(own)
            for c in range(plan.chunks):
                a, e_ = plan.local_rows(c)
                sa, sb = plan.slab_rows(c)
                g_c = gid[a:min(e_, plan.n_local)]
                assert ((g_c >= sa) & (g_c < sb)).all()
                assert torch.equal(g_c - sa, r * plan.cr + torch.arange(g_c.numel()))
            gathered.append(gid)
            assert (plan.dst_local >= 0).all() and (plan.dst_local < plan.n_local).all()
            assert torch.equal(sg.dst[plan.edge_ids] - plan.r0, plan.dst_local)
        gathered = torch.cat(gathered)
        assert torch.unique(gathered).numel() == n and gathered.max() < plan.n_padded
        assert (seen == 1).all()
    # shuffled ids: rows and in-edges balance together
    plans = [RowPartitionPlan(sg.src, sg.dst, n, r, 2) for r in range(2)]
    assert abs(plans[0].num_local_edges - plans[1].num_local_edges) < 0.25 * sg.src.numel()


def test_edge_balanced_bounds():
    sg, _ = _problem()
    n = sg.num_nodes
    deg = torch.bincount(sg.dst, minlength=n)
    for world in (1, 2, 3, 8):
        per, b = block_bounds_by_edges(deg, world)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
        assert per == max(b[i + 1] - b[i] for i in range(world))
        counts = [int(deg[b[i]:b[i + 1]].sum()) for i in range(world)]
        assert sum(counts) == int(deg.sum())
        assert max(counts) - min(counts) <= 2 * int(deg.max())          # within one (hub) row of the ideal cut
        seen = torch.zeros(sg.src.numel(), dtype=torch.int32)
        gathered = []
        for r in range(world):
            plan = RowPartitionPlan(sg.src, sg.dst, n, r, world, chunks=2, balance="edges")
            seen[plan.edge_ids] += 1
            gathered.append(plan.gathered_id(torch.arange(plan.r0, plan.r1)))
            fp = FeaturePartitionPlan(sg.src, sg.dst, n, r, world, balance="edges")
            assert torch.equal(fp.gathered_id(torch.arange(fp.r0, fp.r1)), r * fp.per + torch.arange(fp.n_local))
        gathered = torch.cat(gathered)
        assert (seen == 1).all() and torch.unique(gathered).numel() == n and gathered.max() < plan.n_padded


def _run_world(world, tmp_path, mode="rows", reducer="sum", balance="rows"):
    port = 29500 + (os.getpid() * 7 + world + 11 * len(mode + reducer + balance)) % 2000
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    # spawn, not fork: the parent already runs OpenMP / MKL threads
    mp.start_processes(_worker, args=(world, port, str(tmp_path), mode, reducer, balance), nprocs=world, join=True,
                       start_method="spawn")
    out_ref, grads_ref = _reference(reducer)
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert [p["r0"] for p in parts][0] == 0 and parts[-1]["r1"] == out_ref.shape[0]
    out = torch.cat([p["out"] for p in parts])
    assert torch.allclose(out, out_ref, rtol=1e-5, atol=1e-6)
    for r in range(world):                                   # replicated, all-reduced gradients agree on every rank
        for g, gr in zip(parts[r]["grads"], grads_ref):
            assert torch.allclose(g, gr, rtol=1e-4, atol=1e-7)


def test_two_rank_gloo_matches_single_process_oracle(tmp_path):
    _run_world(2, tmp_path)


def test_three_rank_gloo_uneven_blocks(tmp_path):
    """403 rows over 3 ranks: blocks of different in-edge counts, padded slabs and a chunk count that does not divide the
    block — the exchange must still deliver every source row exactly once; padded rows stay out of the gradients."""
    _run_world(3, tmp_path)


def test_row_partition_max_pool_two_ranks(tmp_path):
    """SAGEConv-pool over the row partition: all-gather of the pooled projections, arg-scatter into a full-height partial,
    reduce-scatter."""
    _run_world(2, tmp_path, reducer="max")


def test_row_partition_edge_balanced_three_ranks(tmp_path):
    _run_world(3, tmp_path, reducer="max", balance="edges")


def test_feature_partition_weighted_sum_two_and_three_ranks(tmp_path):
    _run_world(2, tmp_path, mode="cols")
    _run_world(3, tmp_path, mode="cols", balance="edges")


def test_feature_partition_three_layers_pair_plus_single(tmp_path, monkeypatch):
    """Three layers in the feature partition: layers 0 / 1 run as a pair (transform-first, then aggregate-first: two exchanges
    per direction), layer 2 alone; also the plain order (every layer transform-first) on the same problem."""
    import tests.test_dist_gloo as me
    monkeypatch.setenv("PLAGNN_TEST_DIMS", "24,24,24,12")
    monkeypatch.setattr(me, "DIMS", [24, 24, 24, 12])
    _run_world(2, tmp_path, mode="cols")
    monkeypatch.setenv("PLAGNN_DIST_COLS_ALT", "0")
    _run_world(2, tmp_path, mode="cols")


def test_feature_partition_four_ranks(tmp_path, monkeypatch):
    """World 4 (the one GPU count of the driver's 1 / 2 / 4 / 8 run that the other tests do not cover): feature partition with
    the layers in pairs, widths that are multiples of 4 * world."""
    import tests.test_dist_gloo as me
    monkeypatch.setenv("PLAGNN_TEST_DIMS", "32,32,16")
    monkeypatch.setattr(me, "DIMS", [32, 32, 16])
    _run_world(4, tmp_path, mode="cols", balance="edges")


def test_feature_partition_max_pool_three_ranks(tmp_path):
    _run_world(3, tmp_path, mode="cols", reducer="max")
