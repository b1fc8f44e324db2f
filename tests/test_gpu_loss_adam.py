"""K4 parity: fused loss (value + gradient) vs the reference-generated golden vectors and the oracle;
fused Adam vs torch.optim.Adam; label decision vs the golden vectors."""
import os

import numpy as np
import pytest
import torch

import plagnn_b200 as P
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def test_bce_matches_reference_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_functions.npz"))
    p = torch.tensor(g["loss_p"], device=cuda, requires_grad=True)
    loss = P.multi_loss(p, torch.tensor(g["loss_t"], device=cuda), g["loss_w"])
    loss.backward()
    assert abs(loss.item() - float(g["loss_value"])) <= 1e-6 * abs(float(g["loss_value"]))
    # saturated entries carry gradients up to 1e9 * scale: compare element-wise, relative
    got, want = p.grad.cpu().numpy(), g["loss_grad"]
    assert np.array_equal(got == 0, want == 0)         # clamp masks identical
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=0)


def test_bce_indexed_equals_gather_then_loss(cuda):
    torch.manual_seed(0)
    n, r = 3000, 1100
    prob = torch.sigmoid(torch.randn(n, 12) * 3)
    tgt = (torch.rand(n, 12) < 0.2).float()
    w = np.random.default_rng(0).uniform(1, 30, 12)
    idx = torch.randperm(n)[:r]
    po = prob.clone().requires_grad_(True)
    lo = orc.multi_loss(po[idx.tolist()], tgt[idx.tolist()], w)
    lo.backward()
    pc = P.ops.aligned(prob.to(cuda)).requires_grad_(True)
    lc = P.multi_loss_indexed(pc, tgt.to(cuda), idx.tolist(), w)
    lc.backward()
    assert rel_err(lc.detach(), lo.detach()) < REL_TOL
    assert rel_err(pc.grad, po.grad) < REL_TOL
    assert (pc.grad.cpu()[~torch.isin(torch.arange(n), idx)] == 0).all()


def test_bce_negative_and_out_of_range_indices(cuda):
    """Negative indices count from the end (torch indexing); an index outside [-N, N) — IndexError in the reference —
    must not be skipped silently: the loss comes back NaN."""
    torch.manual_seed(1)
    n, c = 700, 12
    prob = P.ops.aligned(torch.sigmoid(torch.randn(n, c)).to(cuda))
    tgt = (torch.rand(n, c) < 0.3).float().to(cuda)
    w = torch.rand(c, dtype=torch.float64) * 5 + 1
    cw, cwp1 = w.float().to(cuda), (w + 1.0).float().to(cuda)
    pos = torch.tensor([0, 5, 699, 300, 5], device=cuda)                  # a duplicate too
    neg = torch.tensor([-700, -695, -1, -400, -695], device=cuda)
    l_pos, g_pos = P.ops.bce_weighted(prob, tgt, pos, cw, cwp1)
    l_neg, g_neg = P.ops.bce_weighted(prob, tgt, neg, cw, cwp1)
    assert torch.equal(l_pos, l_neg) and torch.equal(g_pos, g_neg) and torch.isfinite(l_pos).all()
    for bad in (n, -n - 1):
        l_bad, _ = P.ops.bce_weighted(prob, tgt, torch.tensor([0, bad, 3], device=cuda), cw, cwp1)
        assert torch.isnan(l_bad).all()
    l_again, _ = P.ops.bce_weighted(prob, tgt, pos, cw, cwp1)              # no state left behind by the failed call
    assert torch.equal(l_again, l_pos)


def test_fused_adam_tracks_torch_adam(cuda):
    torch.manual_seed(1)
    shapes = [(503, 503), (503,), (400, 503), (400,), (12, 100), (12,)]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=5e-5)
    o_mine = P.FusedAdam(mine, lr=5e-5)
    for step in range(6):
        for a, b in zip(ref, mine):
            g = torch.randn(a.shape, generator=torch.Generator().manual_seed(100 + step)) * (10.0 ** (step - 3))
            a.grad, b.grad = g.clone(), g.to(cuda)
        o_ref.step(); o_mine.step()
    for a, b in zip(ref, mine):
        assert rel_err(b, a) < 1e-6
    for a, b in zip(ref, mine):
        sa, sb = o_ref.state[a], o_mine.state[b]
        assert rel_err(sb["exp_avg"], sa["exp_avg"]) < 1e-6
        assert rel_err(sb["exp_avg_sq"], sa["exp_avg_sq"]) < 1e-6


def test_loc_correction_matches_reference_golden(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "reference_functions.npz"))
    pred = P.protein_loc_correction(torch.tensor(g["lc_probs"], device=cuda), float(g["lc_alpha"]))
    assert pred.dtype == torch.float64
    assert np.array_equal(pred.cpu().numpy(), g["lc_pred"])
    aim, cov, acc = P.performances_record(torch.tensor(g["pr_truth"], device=cuda), pred)
    np.testing.assert_allclose([aim, cov, acc], g["pr_out"], rtol=2e-6)


def test_colsum_transpose_padcopy(cuda):
    x = torch.randn(24041, 400, generator=torch.Generator().manual_seed(2))
    xc = P.ops.aligned(x.to(cuda))
    assert rel_err(P.ops.colsum(xc), x.double().sum(0)) < 1e-6
    assert torch.equal(P.ops.transpose(xc).cpu(), x.t())
    y = torch.randn(70, 503).to(cuda)
    ya = P.ops.aligned(y)
    assert ya.stride(0) == 512 and torch.equal(ya, y)
