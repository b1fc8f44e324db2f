"""Row-partition choreography on the CUDA kernels (world = 1, 3 pipelined row chunks) against the oracle, and the
row-range aggregation launch it relies on."""
import numpy as np
import pytest
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth
from plagnn_b200.dist import CudaBackend, DistGCN, PartitionedGraph, RowPartitionPlan, dist_gcn_forward_backward
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def test_spmm_row_ranges_tile_the_full_result(cuda):
    n, e, f = 3000, 90000, 96
    sg = synth.scaled_graph(n, e, seed=3, max_degree=4000)
    g = P.graph((sg.src.numpy(), sg.dst.numpy()), num_nodes=n)
    g.chunk = 64
    g = g.to(cuda)
    csc = g.csc()
    assert csc.counts[1] > 0                               # some rows are split over chunks
    x = ops.aligned(torch.randn(n, f, device=cuda))
    w = sg.weight.to(cuda)
    full = ops.spmm_sum(csc, x, w=w)
    out = ops.alloc(n, f, cuda, zero=True)
    for a, b in ((0, 700), (700, 701), (701, 2500), (2500, n)):
        ops.spmm_sum_rows(csc, ops.plan_range(csc, a, b), x, out, w=w)
    assert torch.equal(out, full)


def test_partitioned_gcn_world1_matches_oracle(cuda):
    n, e, f = 2000, 60000, 48
    sg = synth.scaled_graph(n, e, seed=5, max_degree=3000)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    plan = RowPartitionPlan(sg.src, sg.dst, n, 0, 1, chunks=3)
    pg = PartitionedGraph(plan, sg.weight, P.build_csr, cuda)
    model = DistGCN([f, 40, 24], seed=3)
    ref = orc.GCNSumRef([f, 40, 24])
    with torch.no_grad():
        for lin, w, b in zip(ref.lins, model.weights, model.biases):
            lin.weight.copy_(w); lin.bias.copy_(b)
    model = model.to(cuda)
    h0 = ops.alloc(plan.per, f, cuda, zero=True)
    h0[:n].copy_(x)
    mask = (torch.arange(plan.per, device=cuda) < n).float().unsqueeze(1)
    with torch.cuda.device(cuda):
        out, grads = dist_gcn_forward_backward(model, pg, h0, CudaBackend(pg), None, lambda o: o * mask / n)
    go = orc.OracleGraph(sg.src.numpy(), sg.dst.numpy(), n)
    scale = 1.0 / torch.bincount(sg.dst, minlength=n).clamp(min=1).float()
    out_ref = ref(go, x, sg.weight, scale)
    (0.5 * (out_ref ** 2).sum() / n).backward()
    # gathered numbering is chunk-major: map back to node order before comparing
    assert rel_err(out[:n], out_ref) < REL_TOL
    ref_grads = [t.grad for lin in ref.lins for t in (lin.weight, lin.bias)]
    for g, gr in zip(grads, ref_grads):
        assert rel_err(g, gr) < 2 * REL_TOL
