"""K2 parity: aggregation kernels vs the oracle's C kernels.  max: values and arg bit-exact;
sums and scatter-adds within 1e-5 relative (fp32 accumulation order differs)."""
import numpy as np
import pytest
import torch

import plagnn_b200 as P
from plagnn_b200 import ops
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, random_multigraph, rel_err

pytestmark = pytest.mark.gpu


def make(cuda, n, e, seed, loops=True, chunk=128, **kw):
    src, dst = random_multigraph(n, e, seed, **kw)
    g = P.graph((src, dst), num_nodes=n)
    g.chunk = chunk
    if loops:
        g = g.add_self_loop()
    g = g.to(cuda)
    so, do = (orc.add_self_loop(src.astype(np.int64), dst.astype(np.int64), n) if loops
              else (src.astype(np.int64), dst.astype(np.int64)))
    return g, orc.OracleGraph(so, do, n)


@pytest.mark.parametrize("f", [1, 3, 5, 12, 100, 128, 129, 200, 260, 300, 385, 400, 503, 513, 600, 1024])
def test_spmm_max_forward_bit_exact(cuda, f):
    g, go = make(cuda, 900, 20000, f, loops=False, hubs=3, hub_deg=700, isolated=5)
    x = torch.randn(900, f, generator=torch.Generator().manual_seed(f))
    x = torch.relu(x)                                 # many exact ties at 0, as after fc_pool's ReLU
    out, arg = ops.spmm_max_fwd(g.csc(), x.to(cuda))
    ro, ra = orc.spmm_max_c(*go.csc()[:2], x)
    assert torch.equal(out.cpu(), ro)
    assert torch.equal(arg.cpu(), ra)
    assert (out[:5] == 0).all() and (arg[:5] == -1).all()


def test_spmm_max_chunk_sizes_agree(cuda):
    x = torch.randn(600, 77)
    ref = None
    for chunk in (32, 64, 128, 1024):
        g, go = make(cuda, 600, 30000, 5, chunk=chunk, hubs=2, hub_deg=2500)
        out, arg = ops.spmm_max_fwd(g.csc(), x.to(cuda))
        if ref is None:
            ref = orc.spmm_max_c(*go.csc()[:2], x)
        assert torch.equal(out.cpu(), ref[0]) and torch.equal(arg.cpu(), ref[1])


def make_simple(cuda, n, e, seed):
    """Simple graph (no duplicate pairs) with a few hubs, + self-loops."""
    from plagnn_b200 import synth
    src, dst = synth.powerlaw_edges(n, e, 2.0, seed)
    src, dst = src.numpy().astype(np.int32), dst.numpy().astype(np.int32)
    g = P.graph((src, dst), num_nodes=n)
    g.chunk = 64                                      # small chunks so that some rows are split even on small graphs
    g = g.add_self_loop().to(cuda)
    so, do = orc.add_self_loop(src.astype(np.int64), dst.astype(np.int64), n)
    return g, orc.OracleGraph(so, do, n)


@pytest.mark.parametrize("f", [12, 300, 503])
def test_spmm_max_backward_both_variants(cuda, f):
    g, go = make_simple(cuda, 700, 16000, 21)
    assert not g.has_duplicate_edges() and g.csr().counts[1] > 0      # some rows are split over chunks
    x = torch.relu(torch.randn(700, f))
    out, arg = ops.spmm_max_fwd(g.csc(), x.to(cuda))
    dz = torch.randn(700, f)
    ref = orc.spmm_max_bwd_c(arg.cpu().contiguous(), dz, 700)
    dzc = dz.to(cuda)                                 # contiguous: a different pitch than arg's
    got = ops.spmm_max_bwd(dzc, arg, None, 700)
    assert rel_err(got, ref) < REL_TOL
    got2 = ops.spmm_max_bwd_gather(g.csr(), dzc, arg, None)
    assert rel_err(got2, ref) < REL_TOL
    got3 = ops.spmm_max_bwd_gather(g.csr(), dzc, arg, None)
    assert torch.equal(got2, got3)                    # ordered variant is run-to-run bit-stable
    # folded ReLU mask: gradient only where the pooled value is > 0
    ref_m = orc.spmm_max_bwd_c(arg.cpu().contiguous(), dz * (out.cpu() > 0), 700)
    assert rel_err(ops.spmm_max_bwd(dzc, arg, out, 700), ref_m) < REL_TOL
    assert rel_err(ops.spmm_max_bwd_gather(g.csr(), dzc, arg, out), ref_m) < REL_TOL


@pytest.mark.parametrize("f,weighted,scaled", [(256, True, True), (64, False, False), (503, True, False), (33, False, True)])
def test_spmm_sum_and_transpose(cuda, f, weighted, scaled):
    g, go = make(cuda, 800, 25000, 31, hubs=2, hub_deg=1500, isolated=3)
    x = torch.randn(800, f)
    w = torch.rand(go.num_edges) if weighted else None
    sc = (1.0 / torch.as_tensor(np.maximum(1, np.diff(go.csc()[0])), dtype=torch.float32)) if scaled else None
    indptr, indices, eids = go.csc()
    ref = orc.spmm_sum_c(indptr, indices, x, eids=eids if weighted else None, w=w, scale=sc)
    got = ops.spmm_sum(g.csc(), x.to(cuda), w=None if w is None else w.to(cuda), scale=None if sc is None else sc.to(cuda))
    assert rel_err(got, ref) < REL_TOL
    ip, ix, ei = go.csr()
    ref_t = orc.spmm_sum_c(ip, ix, x, eids=ei if weighted else None, w=w)
    got_t = ops.spmm_sum(g.csr(), x.to(cuda), w=None if w is None else w.to(cuda))
    assert rel_err(got_t, ref_t) < REL_TOL


def test_spmm_sum_fused_epilogue(cuda):
    g, go = make(cuda, 500, 9000, 41)
    x = torch.randn(500, 40)
    b = torch.randn(40)
    indptr, indices, _ = go.csc()
    ref = torch.nn.functional.leaky_relu(orc.spmm_sum_c(indptr, indices, x) + b)
    got = ops.spmm_sum(g.csc(), x.to(cuda), bias=b.to(cuda), act=ops.ACT_LEAKY)
    assert rel_err(got, ref) < REL_TOL
    # dropout epilogue: kept entries scaled by 1/(1-p), mask regenerated identically for backward
    p = 0.25
    d1 = ops.spmm_sum(g.csc(), x.to(cuda), bias=b.to(cuda), act=ops.ACT_LEAKY, dropout_p=p, dropout_seed=9)
    ones = ops.dropout_scale_(ops.aligned(torch.ones(500, 40, device=cuda)).clone(), p, 9)
    assert rel_err(d1, got * ones) < 1e-6
    keep = (ones > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.02


def test_full_size_ppi_aggregation(cuda):
    """N = 24 041, E' = 1 424 041, F = 503 (BASELINE configs[1]): equality with the oracle plus the
    size-independent properties out[v,f] == x[arg[v,f], f] and out >= x[v] (self-loop)."""
    from plagnn_b200 import synth
    prob = synth.ppi_problem(state="inter")
    g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=prob.num_nodes).add_self_loop().to(cuda)
    x = torch.relu(torch.randn(prob.num_nodes, 503, generator=torch.Generator().manual_seed(1)))
    xc = x.to(cuda)
    out, arg = ops.spmm_max_fwd(g.csc(), xc)
    cols = torch.arange(503, device=cuda).expand(prob.num_nodes, 503)
    assert torch.equal(out, xc[arg.long(), cols])
    assert (out >= xc).all()
    s, d = orc.add_self_loop(prob.ppi_row.astype(np.int64), prob.ppi_col.astype(np.int64), prob.num_nodes)
    indptr, indices, _ = orc.coo_to_csc(s, d, prob.num_nodes)
    ro, ra = orc.spmm_max_c(indptr, indices, x)
    assert torch.equal(out.cpu(), ro) and torch.equal(arg.cpu(), ra)


def test_spmm_max_backward_scatter_on_multigraph(cuda):
    """Duplicate edges: the scatter form follows DGL (gradient goes once to the winning source node)."""
    g, go = make(cuda, 500, 12000, 3, hubs=2, hub_deg=600)
    assert g.has_duplicate_edges()
    x = torch.relu(torch.randn(500, 77))
    out, arg = ops.spmm_max_fwd(g.csc(), x.to(cuda))
    dz = torch.randn(500, 77)
    ref = orc.spmm_max_bwd_c(arg.cpu().contiguous(), dz, 500)
    assert rel_err(ops.spmm_max_bwd(dz.to(cuda), arg, None, 500), ref) < REL_TOL
