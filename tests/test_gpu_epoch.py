"""The replayable epoch (plagnn_b200.TrainStep: forward, loss, backward, Adam as four library calls under a CUDA graph)
against the autograd-driven loop the unchanged driver runs (code/train.py:197-205), and the device-side Adam step count."""
import copy
import ctypes

import numpy as np
import pytest
import torch

import plagnn_b200 as P
from plagnn_b200 import _lib, ops, synth
from oracle import plagnn_oracle as orc
from tests.helpers import REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def _problem(cuda, n=1500, e=40000, dims=(3, 60, 60)):
    prob = synth.ppi_problem(n, e, "normal", 70, feat_dims=dims)
    ids = list(range(n))
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(cuda)
    torch.manual_seed(3)
    m = P.GNN32(sum(dims), 400, 300, 200, 100, 12).to(cuda)
    w = orc.weight_cal(prob.loc)
    idx = torch.as_tensor(prob.labelled[::2], device=cuda)
    return prob, g, m, w, idx


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_matches_the_autograd_loop(cuda, use_graph):
    prob, g, m0, w, idx = _problem(cuda)
    x, y = g.ndata["feat"], g.ndata["loc"]
    ma, mb = copy.deepcopy(m0), copy.deepcopy(m0)
    opt = P.FusedAdam(ma.parameters(), lr=5e-5)
    ts = P.TrainStep(mb, g, x, y, idx, w, lr=5e-5, use_graph=use_graph)
    for p0, pb in zip(m0.parameters(), mb.parameters()):
        assert torch.equal(p0, pb)                       # the warm-up before the capture leaves no trace
    assert int(ts.step_count.item()) == 0
    for ep in range(4):
        opt.zero_grad()
        la = ma(g, x)
        loss_a = P.multi_loss_indexed(la, y, idx, w)
        loss_a.backward()
        opt.step()
        ts.step()
        assert abs(ts.loss.item() - loss_a.item()) <= REL_TOL * abs(loss_a.item())
        assert rel_err(ts.logits, la) < (REL_TOL if ep == 0 else 1e-4)      # unordered fp32 reductions in the max backward
    assert int(ts.step_count.item()) == 4 and ts.epochs == 4
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert (pa - pb).abs().max().item() <= 2.5 * 5e-5 * 4               # Adam moves an entry by at most ~lr per step
        assert rel_err(pb, pa) < 1e-3
    if use_graph:
        assert ts.launches_per_epoch and ts.launches_per_epoch > 40


def test_adam_with_the_step_count_on_the_device(cuda):
    lib = _lib.load()
    torch.manual_seed(0)
    shapes = [(300, 200), (17,), (64, 3)]
    pa = [torch.randn(s, device=cuda) for s in shapes]
    pb = [p.clone() for p in pa]
    st = {k: [torch.zeros_like(p) for p in pa] for k in ("ma", "va", "mb", "vb")}
    grads = [torch.empty_like(p) for p in pa]

    def table(ps, ms, vs):
        return torch.tensor([(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
                             for p, g, m, v in zip(ps, grads, ms, vs)], dtype=torch.int64).to(cuda)

    ta, tb = table(pa, st["ma"], st["va"]), table(pb, st["mb"], st["vb"])
    count = torch.zeros(1, dtype=torch.int64, device=cuda)
    scal = torch.zeros(4, device=cuda)
    mx = max(p.numel() for p in pa)
    for step in range(1, 7):
        for g in grads:
            g.copy_(torch.randn_like(g))
        ops.adam_multi(ta, len(pa), mx, 1e-3, 0.9, 0.999, 1e-8, step)
        _lib.check(lib.plagnn_adam_multi_devstep(tb.data_ptr(), len(pb), mx, 1e-3, 0.9, 0.999, 1e-8, count.data_ptr(),
                                                 scal.data_ptr(), ops._stream()), "adam_multi_devstep")
    assert int(count.item()) == 6
    for a, b in zip(pa, pb):
        assert rel_err(b, a) < 1e-6
    for a, b in zip(st["va"], st["vb"]):
        assert torch.equal(a, b)                          # the second moment does not depend on the bias corrections


def test_loss_gradient_times_incoming_gradient(cuda):
    """loss.backward() multiplies the saved loss gradient by autograd's incoming scalar inside the library."""
    torch.manual_seed(1)
    p = torch.rand(500, 12, device=cuda).clamp(1e-4, 1 - 1e-4).requires_grad_()
    t = (torch.rand(500, 12, device=cuda) > 0.7).float()
    w = np.linspace(0.5, 6.0, 12)
    (P.multi_loss(p, t, w) * 3.0).backward()
    g3 = p.grad.clone()
    p.grad = None
    P.multi_loss(p, t, w).backward()
    assert rel_err(g3, 3.0 * p.grad) < 1e-6
