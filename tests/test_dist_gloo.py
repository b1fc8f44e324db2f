"""N > 1 host logic on CPU: world_size-2 and -3 gloo runs of the row-partitioned GCN choreography with a CPU compute backend
(oracle kernels), compared with the single-process oracle model.  No CUDA involved."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plagnn_oracle as orc
from plagnn_b200 import synth
from plagnn_b200.dist import DistGCN, RowPartitionPlan, block_bounds, dist_gcn_forward_backward


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

    def spmm_rows(self, csx, key, a, b, x, out, w, scale, bias, act):
        r = orc.spmm_sum_c(csx.indptr[a:b + 1], csx.indices, x.contiguous(), eids=csx.eids if w is not None else None, w=w,
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
    def __init__(self, plan, weight):
        self.plan = plan
        self.csc = _Csx(plan.dst_local, plan.src_gathered, plan.per)
        self.csr_t = _Csx(plan.src_gathered, plan.dst_local, plan.n_padded)
        self.edge_weight = weight[plan.edge_ids]
        self.scale = plan.scale_local


def _problem(n=403, e=6000, f=24):
    sg = synth.scaled_graph(n, e, seed=5, max_degree=200)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    return sg, x


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sg, x = _problem()
    n, f = sg.num_nodes, x.shape[1]
    plan = RowPartitionPlan(sg.src, sg.dst, n, rank, world, chunks=3)
    pg = _PG(plan, sg.weight)
    model = DistGCN([f, 16, 8], seed=3)
    h0 = torch.zeros(plan.per, f)
    h0[:plan.n_local] = x[plan.r0:plan.r1]
    mask = (torch.arange(plan.per) < plan.n_local).float().unsqueeze(1)
    out, grads = dist_gcn_forward_backward(model, pg, h0, CpuBackend(), None, lambda o: o * mask / n)
    torch.save({"out": out[:plan.n_local], "grads": grads, "r0": plan.r0, "r1": plan.r1},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def _reference():
    sg, x = _problem()
    n, f = sg.num_nodes, x.shape[1]
    g = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    model = DistGCN([f, 16, 8], seed=3)
    ref = orc.GCNSumRef([f, 16, 8])
    with torch.no_grad():
        for lin, w, b in zip(ref.lins, model.weights, model.biases):
            lin.weight.copy_(w); lin.bias.copy_(b)
    scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
    out = ref(g, x, sg.weight, scale)
    (0.5 * (out ** 2).sum() / n).backward()
    grads = [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]
    return out.detach(), grads


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


def _run_world(world, tmp_path):
    port = 29500 + (os.getpid() * 7 + world) % 2000
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    # spawn, not fork: the parent already runs OpenMP / MKL threads
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    out_ref, grads_ref = _reference()
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
    block — the exchange must still deliver every source row exactly once."""
    _run_world(3, tmp_path)
