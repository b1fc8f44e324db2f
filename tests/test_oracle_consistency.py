"""The DGL part of the oracle cannot be pinned against DGL (not installable here); instead three
independent formulations of copy_u+max must agree, the custom backward must pass gradcheck, and the
structural invariants of the max reducer must hold."""
import numpy as np
import pytest
import torch

from oracle import plagnn_oracle as orc
from tests.helpers import random_multigraph


@pytest.mark.parametrize("n,e,f", [(50, 400, 7), (120, 900, 33), (64, 64, 5)])
def test_three_formulations_agree(n, e, f):
    src, dst = random_multigraph(n, e, 1, isolated=3)
    indptr, indices, _ = orc.coo_to_csc(src, dst, n)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(0))
    x[x.abs() < 0.3] = 0.0           # plenty of exact ties, like post-ReLU activations
    oc, ac = orc.spmm_max_c(indptr, indices, x)
    ol, al = orc.spmm_max_loop(indptr, indices, x)
    od = orc.spmm_max_dense(src, dst, n, x)
    assert torch.equal(oc, ol) and torch.equal(oc, od)
    assert torch.equal(ac, al)       # first maximum in in-edge order in both
    assert (oc[:3] == 0).all() and (ac[:3] == -1).all()   # rows without in-edges -> 0


def test_first_max_wins_in_edge_order():
    # two sources with the same value: the one that comes first in CSC (stable by edge id) wins
    src = np.array([4, 2, 3], dtype=np.int32)
    dst = np.array([0, 0, 0], dtype=np.int32)
    indptr, indices, eids = orc.coo_to_csc(src, dst, 5)
    assert list(indices) == [4, 2, 3] and list(eids) == [0, 1, 2]
    x = torch.zeros(5, 2)
    x[2, 0] = x[4, 0] = 1.0
    x[3, 1] = x[2, 1] = 2.0
    _, arg = orc.spmm_max_c(indptr, indices, x)
    assert arg[0].tolist() == [4, 2]


def test_backward_c_matches_index_put_and_gradcheck():
    n, e, f = 40, 300, 6
    src, dst = random_multigraph(n, e, 3)
    indptr, indices, _ = orc.coo_to_csc(src, dst, n)
    x = torch.randn(n, f, requires_grad=True)
    w = torch.randn(n, f)
    (orc.spmm_max(x, indptr, indices, True) * w).sum().backward()
    g_c = x.grad.clone()
    x.grad = None
    (orc.spmm_max(x, indptr, indices, False) * w).sum().backward()
    assert torch.allclose(g_c, x.grad, rtol=1e-6, atol=1e-6)
    xd = torch.randn(n, f, dtype=torch.float64, requires_grad=True)   # distinct values almost surely
    assert torch.autograd.gradcheck(lambda t: orc.spmm_max(t, indptr, indices, False), (xd,), eps=1e-6, atol=1e-5)


def test_invariants_edge_permutation_and_non_argmax_removal():
    n, e, f = 60, 500, 9
    src, dst = random_multigraph(n, e, 4)
    x = torch.randn(n, f)
    indptr, indices, _ = orc.coo_to_csc(src, dst, n)
    out, arg = orc.spmm_max_c(indptr, indices, x)
    perm = np.random.default_rng(0).permutation(e)
    ip2, ix2, _ = orc.coo_to_csc(src[perm], dst[perm], n)
    out2, _ = orc.spmm_max_c(ip2, ix2, x)
    assert torch.equal(out, out2)                     # values do not depend on edge order
    # drop one in-edge of row v that is not an arg-max for any feature: nothing changes
    for v in range(n):
        b, en = indptr[v], indptr[v + 1]
        cand = [k for k in range(b, en) if indices[k] not in set(arg[v].tolist())]
        if cand:
            keep = np.ones(len(indices), bool)
            keep[cand[0]] = False
            s2, d2 = indices[keep], np.repeat(np.arange(n), np.diff(indptr))[keep]
            ip3, ix3, _ = orc.coo_to_csc(s2, d2, n)
            out3, arg3 = orc.spmm_max_c(ip3, ix3, x)
            assert torch.equal(out, out3) and torch.equal(arg, arg3)
            break


def test_csc_is_stable_sort_with_self_loops_last():
    n = 30
    src, dst = random_multigraph(n, 200, 6)
    s, d = orc.add_self_loop(src, dst, n)
    indptr, indices, eids = orc.coo_to_csc(s, d, n)
    assert indptr[0] == 0 and indptr[-1] == len(s)
    for v in range(n):
        seg = eids[indptr[v]:indptr[v + 1]]
        assert (np.diff(seg) > 0).all()               # stable: edge ids ascending inside a row
        assert seg[-1] == 200 + v and indices[indptr[v + 1] - 1] == v   # the self-loop is last
        assert (d[seg] == v).all() and (s[seg] == indices[indptr[v]:indptr[v + 1]]).all()


def test_sum_kernel_and_transpose_are_adjoint():
    n, e, f = 45, 350, 8
    src, dst = random_multigraph(n, e, 8)
    g = orc.OracleGraph(src, dst, n)
    w = torch.rand(e)
    scale = torch.rand(n) + 0.5
    x = torch.randn(n, f, requires_grad=True)
    y = orc.spmm_sum(x, g, w, scale)
    a = torch.zeros(n, n)
    a.index_put_((torch.as_tensor(dst, dtype=torch.long), torch.as_tensor(src, dtype=torch.long)), w, accumulate=True)
    ref = scale.unsqueeze(1) * (a @ x.detach())
    assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
    r = torch.randn(n, f)
    (y * r).sum().backward()
    assert torch.allclose(x.grad, a.t() @ (scale.unsqueeze(1) * r), rtol=1e-5, atol=1e-5)


def test_sage_layer_float64_gradcheck():
    n, e = 12, 40
    src, dst = random_multigraph(n, e, 9)
    g = orc.OracleGraph(*orc.add_self_loop(src.astype(np.int64), dst.astype(np.int64), n), n)
    layer = orc.SAGEConvPoolRef(5, 4, "pool", use_c=False).double()
    x = torch.randn(n, 5, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: layer(g, t), (x,), eps=1e-6, atol=1e-5)


def _small_problem(n=600, e=12000, dims=(3, 40, 40), seed=70):
    import plagnn_b200 as P
    from plagnn_b200 import synth
    prob = synth.ppi_problem(n, e, "normal", seed, feat_dims=dims)
    ids = list(range(n))
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    torch.manual_seed(seed)
    mo = orc.GNN32Ref(sum(dims), 400, 300, 200, 100, 12)
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in prob.labelled[::2]]
    return prob, go, mo, w, idx


def test_shared_decisions_reproduce_the_plain_forward_and_backward():
    """forward_with_decisions fed the model's OWN decisions is the model: same output, same gradients, no differing entry."""
    prob, go, mo, w, idx = _small_problem()
    x, y = go.ndata["feat"], go.ndata["loc"]
    out = mo(go, x)
    orc.multi_loss(out[idx], y[idx], w).backward()
    plain = [p.grad.clone() for p in mo.parameters()]
    mo.zero_grad()
    out2, rep = orc.forward_with_decisions(mo, go, x, orc.own_decisions(mo, go, x))
    orc.multi_loss(out2[idx], y[idx], w).backward()
    assert rep["arg_diff"] == rep["pool_diff"] == rep["act_diff"] == rep["bad_edges"] == 0 and rep["near_ties_only"]
    assert torch.allclose(out, out2, rtol=0, atol=1e-7)
    for a, p in zip(plain, mo.parameters()):
        assert (a - p.grad).abs().max() <= 2e-6 * a.abs().max()


def test_fp32_vs_float64_gradient_gap_is_the_discrete_decisions():
    """The cause of the 1e-4-level gradient differences between two correct implementations, shown on the CPU: the fp32
    and the float64 oracle differ by far more than 1e-5 on some gradients, and agree to 1e-5 on all of them once the
    float64 model is evaluated on the fp32 model's decisions — every one of which is a near-tie in float64."""
    import copy
    from tests.helpers import rel_err
    prob, go, mo, w, idx = _small_problem(n=2000, e=60000, dims=(3, 250, 250))   # here: 7 of 2.4 M arg-max entries differ, plain gap 4.7e-5
    x, y = go.ndata["feat"], go.ndata["loc"]
    md = copy.deepcopy(mo).double()
    orc.multi_loss(mo(go, x)[idx], y[idx], w).backward()
    orc.multi_loss(md(go, x.double())[idx], y[idx].double(), w).backward()
    plain = max(rel_err(a.grad, b.grad) for a, b in zip(mo.parameters(), md.parameters()))
    md.zero_grad()
    out, rep = orc.forward_with_decisions(md, go, x.double(), orc.own_decisions(mo, go, x))
    orc.multi_loss(out[idx], y[idx].double(), w).backward()
    shared = max(rel_err(a.grad, b.grad) for a, b in zip(mo.parameters(), md.parameters()))
    assert rep["near_ties_only"], rep
    assert shared < 1e-5, (shared, plain, rep)
    assert shared <= plain
