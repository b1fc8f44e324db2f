"""End-to-end parity of the drop-in model against the oracle on the same inputs and weights:
fp32 logits, gradients and per-epoch loss within 1e-5 relative, identical predicted labels
(BASELINE.json north_star), plus the reference-loop golden fixture."""
import os

import numpy as np
import pytest
import torch
from scipy.sparse import coo_matrix

import plagnn_b200 as P
from plagnn_b200 import synth
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, copy_params, rel_err
from tests.test_oracle_train_loop import load_problem, run_loop

pytestmark = pytest.mark.gpu


def build_pair(cuda, n=2000, e=60000, seed=70, dims=(3, 250, 250)):
    prob = synth.ppi_problem(n, e, "normal", seed, feat_dims=dims)
    ids = list(range(n))
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(cuda)
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    torch.manual_seed(seed)
    mo = orc.GNN32Ref(sum(dims), 400, 300, 200, 100, 12)
    m = P.GNN32(sum(dims), 400, 300, 200, 100, 12)
    copy_params(m, mo)
    return prob, g, go, m.to(cuda), mo


def epoch_cuda(model, opt, g, features, labels, train_index, i_weight):
    opt.zero_grad()
    model.train()
    logits = model(g, features)
    loss = P.multi_loss(logits[train_index], labels[train_index], i_weight)
    loss.backward()
    opt.step()
    return logits, loss


def test_same_init_as_oracle_given_the_seed():
    torch.manual_seed(5)
    a = P.GNN32(43, 400, 300, 200, 100, 12)
    torch.manual_seed(5)
    b = orc.GNN32Ref(43, 400, 300, 200, 100, 12)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert sum(p.numel() for p in P.GNN32(503, 400, 300, 200, 100, 12).parameters()) == 1288824


@pytest.mark.parametrize("backend", ["simt", "tcgen05"])
def test_forward_backward_parity(cuda, backend, monkeypatch):
    from plagnn_b200 import ops
    forced = ops.GEMM_SIMT if backend == "simt" else ops.GEMM_AUTO
    orig = ops.gemm
    monkeypatch.setattr(ops, "gemm", lambda *a, **k: orig(*a, **{**k, "backend": forced}))
    prob, g, go, m, mo = build_pair(cuda)
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in prob.labelled[: len(prob.labelled) * 9 // 10]]
    lo = mo(go, go.ndata["feat"])
    loss_o = orc.multi_loss(lo[idx], go.ndata["loc"][idx], w)
    loss_o.backward()
    lc = m(g, g.ndata["feat"])
    loss_c = P.multi_loss(lc[idx], g.ndata["loc"][idx], w)
    loss_c.backward()
    assert rel_err(lc, lo) < REL_TOL, "logits"
    assert abs(loss_c.item() - loss_o.item()) <= REL_TOL * abs(loss_o.item()), "loss"
    for (name, pc), (_, po) in zip(m.named_parameters(), mo.named_parameters()):
        assert rel_err(pc.grad, po.grad) < 2 * REL_TOL, f"grad {name}: {rel_err(pc.grad, po.grad):.3e}"
    pred_c = P.protein_loc_correction(lc, 0.1).cpu()
    pred_o = orc.protein_loc_correction(lo.detach(), 0.1)
    assert (pred_c != pred_o).float().mean().item() < 1e-4          # identical labels up to fp32 near-ties
    pred_same_input = P.protein_loc_correction(lo.detach().to(cuda), 0.1).cpu()
    assert torch.equal(pred_same_input, pred_o)                     # same probabilities -> identical labels


def test_fused_function_equals_layerwise_path(cuda):
    prob, g, go, m, mo = build_pair(cuda, n=800, e=16000, dims=(3, 20, 20))
    x = g.ndata["feat"]
    a = m(g, x)
    b = m.forward_layerwise(g, x)
    assert rel_err(a, b) < 1e-6
    ga = torch.autograd.grad(a.sum(), list(m.parameters()))
    gb = torch.autograd.grad(b.sum(), list(m.parameters()))
    for u, v in zip(ga, gb):
        assert rel_err(u, v) < 1e-5


def test_deterministic_backward_option(cuda):
    from plagnn_b200 import nn as pnn
    prob, g, go, m, mo = build_pair(cuda, n=800, e=16000, dims=(3, 20, 20))
    x = g.ndata["feat"]
    pnn.DETERMINISTIC_BACKWARD = True
    try:
        g1 = torch.autograd.grad(m(g, x).square().sum(), list(m.parameters()))
        g2 = torch.autograd.grad(m(g, x).square().sum(), list(m.parameters()))
    finally:
        pnn.DETERMINISTIC_BACKWARD = False
    g3 = torch.autograd.grad(m(g, x).square().sum(), list(m.parameters()))
    for u, v, w in zip(g1, g2, g3):
        assert torch.equal(u, v)
        assert rel_err(w, u) < 1e-5


def test_three_epochs_track_the_oracle(cuda):
    prob, g, go, m, mo = build_pair(cuda, n=1500, e=40000)
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in prob.labelled[::2]]
    oo = torch.optim.Adam(mo.parameters(), lr=5e-5)
    oc = P.FusedAdam(m.parameters(), lr=5e-5)
    for _ in range(3):
        lo, loss_o = orc.train_epoch(mo, oo, go, go.ndata["feat"], go.ndata["loc"], idx, w)
        lc, loss_c = epoch_cuda(m, oc, g, g.ndata["feat"], g.ndata["loc"], idx, w)
        assert abs(loss_c.item() - loss_o.item()) <= REL_TOL * abs(loss_o.item())
        assert rel_err(lc, lo) < 2 * REL_TOL


def test_reference_loop_golden_fixture(cuda, golden_dir):
    """The fixture was produced by the reference's unchanged train.train(); reproduce it on the GPU with the
    drop-in model / graph / loss / Adam under the same seeds."""
    z, ppi, n = load_problem(golden_dir)
    uniprot = list(range(n))
    tl, vl, lg = run_loop(
        z, lambda: P.create_graph(ppi, z["ecc"], z["gcn"], coo_matrix(z["loc"]), z["expr"], uniprot).to(cuda),
        lambda f: P.GNN32(f, 400, 300, 200, 100, 12), lambda p, lr: P.FusedAdam(p, lr=lr), cuda,
        epoch_cuda, P.multi_loss)
    np.testing.assert_allclose(tl, z["train_loss"], rtol=REL_TOL)
    np.testing.assert_allclose(vl, z["val_loss"], rtol=REL_TOL)
    assert rel_err(lg, z["logits"]) < 2 * REL_TOL


def test_sage_mean_and_gcn_sum_family(cuda):
    n, e = 1200, 30000
    src, dst = synth.powerlaw_edges(n, e, 2.2, 3)
    w = torch.rand(e)
    g = P.graph((src.numpy(), dst.numpy()), num_nodes=n).to(cuda)
    go = orc.OracleGraph(src.numpy(), dst.numpy(), n)
    scale = 1.0 / torch.bincount(dst, minlength=n).clamp(min=1).float()
    torch.manual_seed(0)
    mo = orc.GCNSumRef([64, 96, 32])
    m = P.GCN([64, 96, 32])
    with torch.no_grad():
        for layer, lin in zip(m.layers, mo.lins):
            layer.weight.copy_(lin.weight); layer.bias.copy_(lin.bias)
    m = m.to(cuda)
    x = torch.randn(n, 64)
    yo = mo(go, x, w, scale)
    yc = m(g, x.to(cuda), w.to(cuda), scale.to(cuda))
    assert rel_err(yc, yo) < REL_TOL
    r = torch.randn(n, 32)
    (yo * r).sum().backward()
    (yc * r.to(cuda)).sum().backward()
    for layer, lin in zip(m.layers, mo.lins):
        assert rel_err(layer.weight.grad, lin.weight.grad) < 2 * REL_TOL
        assert rel_err(layer.bias.grad, lin.bias.grad) < 2 * REL_TOL
