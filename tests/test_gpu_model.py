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


def assert_gradients_match_on_shared_decisions(m, g, mo, go, idx, w, tag, tol=REL_TOL):
    """Gradient bar of the north_star (1e-5 relative, every tensor) with the oracle evaluated on the CUDA path's own
    discrete decisions; every decision that differs from the oracle's own must be a near-tie in the oracle's numbers and
    every arg id a real in-edge (oracle.forward_with_decisions)."""
    dec = m.discrete_decisions(g, g.ndata["feat"])
    mo.zero_grad()
    lo, rep = orc.forward_with_decisions(mo, go, go.ndata["feat"], dec)
    orc.multi_loss(lo[idx], go.ndata["loc"][idx], w).backward()
    lines = [f"{tag}: decisions differing from the fp32 oracle's own: arg {rep['arg_diff']} (gap {rep['arg_gap']:.1e}), "
             f"pool sign {rep['pool_diff']} ({rep['pool_gap']:.1e}), leaky branch {rep['act_diff']} ({rep['act_gap']:.1e}) "
             f"of {rep['entries']} pooled entries; arg ids that are not in-edges: {rep['bad_edges']}"]
    bad = []
    for (name, pc), po in zip(m.named_parameters(), mo.parameters()):
        e = rel_err(pc.grad, po.grad)
        lines.append(f"{name:24s} cuda-vs-oracle32 (shared decisions) {e:.2e}")
        if not e <= tol:
            bad.append((name, e))
    print("\n".join(lines))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/grad_parity_shared_{tag}.txt", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    assert rep["near_ties_only"], rep
    assert not bad, f"gradient parity (shared decisions, bar {tol:g}) failed: {bad}"


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
    m.engine = "python" if backend == "simt" else "c"       # python-orchestrated FFMA path vs the C engine (tcgen05)
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
    # Gradients, 1e-5 on every tensor.  GNN32 is piecewise linear and max-pool / leaky_relu are discontinuous in their
    # choice of piece: two fp32 pipelines that agree to 2e-6 on every activation still resolve a handful of near-ties
    # (relative gap < 1e-6) differently, and ONE re-routed entry moves a weight gradient by ~1e-4 (CPU-side proof with
    # the fp32 and float64 oracles: tests/test_oracle_consistency.py).  So the oracle's backward runs on the CUDA path's
    # decisions, each checked to be a near-tie in the oracle's own numbers; what is compared is then the arithmetic.
    plain = [f"{name:24s} cuda-vs-oracle32 (own decisions) {rel_err(pc.grad, po.grad):.2e}"
             for (name, pc), po in zip(m.named_parameters(), mo.parameters())]
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/grad_parity_{backend}.txt", "w") as fh:
        fh.write(f"logits rel err {rel_err(lc, lo):.3e}  loss rel err {abs(loss_c.item() - loss_o.item()) / abs(loss_o.item()):.3e}\n"
                 + "\n".join(plain) + "\n")
    assert_gradients_match_on_shared_decisions(m, g, mo, go, idx, w, backend)
    pred_c = P.protein_loc_correction(lc, 0.1).cpu()
    pred_o = orc.protein_loc_correction(lo.detach(), 0.1)
    assert (pred_c != pred_o).float().mean().item() < 1e-4          # identical labels up to fp32 near-ties
    pred_same_input = P.protein_loc_correction(lo.detach().to(cuda), 0.1).cpu()
    assert torch.equal(pred_same_input, pred_o)                     # same probabilities -> identical labels


def test_engine_paths_agree(cuda):
    """C engine (two whole-network calls), python-orchestrated fused Function and per-layer Functions."""
    prob, g, go, m, mo = build_pair(cuda, n=800, e=16000, dims=(3, 20, 20))
    x = g.ndata["feat"]
    m.engine = "c"
    a = m(g, x)
    m.engine = "python"
    a2 = m(g, x)
    b = m.forward_layerwise(g, x)
    assert torch.equal(a, a2)                           # same kernels, same order -> bit-identical
    assert rel_err(a, b) < 1e-6
    ga = torch.autograd.grad(a.sum(), list(m.parameters()))
    ga2 = torch.autograd.grad(a2.sum(), list(m.parameters()))
    gb = torch.autograd.grad(b.sum(), list(m.parameters()))
    for u, u2, v in zip(ga, ga2, gb):
        assert rel_err(u, u2) < 1e-5                    # fp32 reductions of the scatter are unordered
        assert rel_err(u, v) < 1e-5
    # a second forward before the first backward must not clobber the first one's saved state
    m.engine = "c"
    y1 = m(g, x)
    y2 = m(g, x * 0.5)
    g1 = torch.autograd.grad(y1.sum(), list(m.parameters()))
    for u, v in zip(ga, g1):
        assert rel_err(u, v) < 1e-5
    with torch.no_grad():
        assert torch.equal(m(g, x), y1)


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
    for ep in range(3):
        lo, loss_o = orc.train_epoch(mo, oo, go, go.ndata["feat"], go.ndata["loc"], idx, w)
        lc, loss_c = epoch_cuda(m, oc, g, g.ndata["feat"], g.ndata["loc"], idx, w)
        # per-epoch loss within 1e-5 (north_star).  Logits of later epochs see the parameters after Adam, whose
        # first steps move every weight by ~lr*sign(g): noise-level gradients flip sign between any two fp32
        # implementations, so the outputs drift apart at the 1e-5..1e-4 level (epoch 1 is held to 1e-5).
        assert abs(loss_c.item() - loss_o.item()) <= REL_TOL * abs(loss_o.item())
        assert rel_err(lc, lo) < (REL_TOL if ep == 0 else 1e-4)


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


def test_concurrent_models_on_streams_match_sequential(cuda):
    """Two models of the fold/seed sweep trained at the same time on two CUDA streams (one arena each) end where the same
    two models end when trained one after the other (ordered max-pool backward, so the comparison is exact)."""
    import copy
    from plagnn_b200 import nn as pnn
    prob, g, go, m0, mo = build_pair(cuda, n=1500, e=40000, dims=(3, 60, 60))
    w = orc.weight_cal(prob.loc)
    idx = torch.as_tensor(prob.labelled[::2], device=cuda)
    x, y = g.ndata["feat"], g.ndata["loc"]
    torch.manual_seed(5)
    m1 = P.GNN32(123, 400, 300, 200, 100, 12).to(cuda)
    seq = [copy.deepcopy(m0), copy.deepcopy(m1)]
    par = [copy.deepcopy(m0), copy.deepcopy(m1)]

    def step(model, opt):
        opt.zero_grad()
        loss = P.multi_loss_indexed(model(g, x), y, idx, w)
        loss.backward()
        opt.step()

    pnn.DETERMINISTIC_BACKWARD = True
    try:
        for m in seq:
            o = P.FusedAdam(m.parameters(), lr=5e-5)
            for _ in range(3):
                step(m, o)
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(device=cuda) for _ in par]
        opts = [P.FusedAdam(m.parameters(), lr=5e-5) for m in par]
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for _ in range(3):
            for m, o, s in zip(par, opts, streams):
                with torch.cuda.stream(s):
                    step(m, o)
        torch.cuda.synchronize()
    finally:
        pnn.DETERMINISTIC_BACKWARD = False
    for a, b in zip(seq, par):
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.equal(pa, pb)


GRAD_TOL_TCGEN05_FULL_SIZE = 2e-5


def _full_size_ppi_epoch_parity(cuda, grad_tol, tag):
    import copy
    prob = synth.ppi_problem(state="inter")
    n = prob.num_nodes
    ids = list(range(n))
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(cuda)
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    torch.manual_seed(70)
    mo = orc.GNN32Ref(503, 400, 300, 200, 100, 12)
    m = P.GNN32(503, 400, 300, 200, 100, 12)
    copy_params(m, mo)
    m = m.to(cuda)
    md = copy.deepcopy(mo).double()
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in prob.labelled[::2]]
    lo = mo(go, go.ndata["feat"])
    loss_o = orc.multi_loss(lo[idx], go.ndata["loc"][idx], w)
    ld = md(go, go.ndata["feat"].double())
    loss_d = orc.multi_loss(ld[idx], go.ndata["loc"][idx].double(), w)
    lc = m(g, g.ndata["feat"])
    loss_c = P.multi_loss_indexed(lc, g.ndata["loc"], torch.as_tensor(idx, device=cuda), w)
    assert rel_err(lc, lo) < REL_TOL and rel_err(lc, ld) < REL_TOL
    assert abs(loss_c.item() - loss_o.item()) <= REL_TOL * abs(loss_o.item())
    assert abs(loss_c.item() - loss_d.item()) <= REL_TOL * abs(loss_d.item())
    pred_c = P.protein_loc_correction(lc, 0.1).cpu()
    pred_o = orc.protein_loc_correction(lo.detach(), 0.1)
    assert (pred_c != pred_o).double().mean().item() <= 1e-4
    loss_c.backward()
    assert_gradients_match_on_shared_decisions(m, g, mo, go, idx, w, tag, tol=grad_tol)


def test_full_size_ppi_epoch_parity(cuda):
    """One epoch at BASELINE.json's full size (N = 24 041, E = 1.4 M + self-loops, F = 503) against the oracle on the DEFAULT
    path (tcgen05, 3 x TF32, parity mode: two accumulation chains per tile, weight-gradient chains of 24 k-blocks): logits and
    loss within 1e-5 relative, predicted localisation labels identical up to fp32 near-ties (<= 1e-4 of the entries), and all
    19 gradient tensors within 1e-5 of the fp32 oracle on shared decisions (measured: <= 8.8e-6).  Until the last session of
    round 2 the default path sat at 2.2e-6 .. 1.5e-5 here: the tensor core truncates when it adds into its fp32 accumulator, the
    forward products (K = 503 / 1006, one chain each) carried a bias of ~2e-6 that the gradient sums over 24 041 rows amplify
    ~3.5 x.  Halving the chains halves the bias; the single-chain kernels remain as PLAGNN_GEMM_PARITY=0
    (test_full_size_fast_gemm_mode)."""
    _full_size_ppi_epoch_parity(cuda, REL_TOL, "full_size")


@pytest.mark.skipif(os.environ.get("PLAGNN_STRICT_INNER") != "2", reason="runs inside test_full_size_fast_gemm_mode")
def test_full_size_fast_inner(cuda):
    _full_size_ppi_epoch_parity(cuda, GRAD_TOL_TCGEN05_FULL_SIZE, "full_size_fast_gemm")


def test_full_size_fast_gemm_mode(cuda):
    """The same full-size epoch with PLAGNN_GEMM_PARITY=0 (one accumulation chain per tile: 460-465 instead of 457 epochs/s): gradients within 2e-5 (measured 2.2e-6 .. 1.5e-5, 8 of 19 tensors above 1e-5), everything else at the bar.
    Own process: the engine's arena is sized for the mode's split-K workspace."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLAGNN_GEMM_PARITY="0", PLAGNN_STRICT_INNER="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu",
                        "-k", "test_full_size_fast_inner"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.skipif(os.environ.get("PLAGNN_STRICT_INNER") != "1", reason="runs inside test_full_size_parity_exact_fp32_gemm")
def test_full_size_strict_inner(cuda):
    _full_size_ppi_epoch_parity(cuda, REL_TOL, "full_size_fp32_gemm")


def test_full_size_parity_exact_fp32_gemm(cuda):
    """The same full-size epoch with every dense product on the FFMA kernel (PLAGNN_GEMM=simt, read once per process, hence
    the child process): all 19 gradient tensors within 1e-5 of the fp32 oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLAGNN_GEMM="simt", PLAGNN_STRICT_INNER="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu",
                        "-k", "test_full_size_strict_inner"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]


def test_model_on_second_gpu_while_device_zero_is_current(cuda):
    """The reference's `-d cuda:1` (code/main_normal.py:30,66): graph and model moved to GPU 1 while the process's current
    device stays 0.  Every wrapper must run its kernels (and take its stream) on the device that holds the tensors."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    dev1 = torch.device("cuda:1")
    assert torch.cuda.current_device() == 0
    prob, g0, go, m0, mo = build_pair(cuda, n=800, e=16000, dims=(3, 20, 20))
    ids = list(range(prob.num_nodes))
    g1 = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(dev1)
    import copy
    m1 = copy.deepcopy(m0).to(dev1)
    w = orc.weight_cal(prob.loc)
    idx = [int(i) for i in prob.labelled[::2]]
    outs = []
    for g, m in ((g0, m0), (g1, m1)):
        opt = P.FusedAdam(m.parameters(), lr=5e-5)
        lc, loss = epoch_cuda(m, opt, g, g.ndata["feat"], g.ndata["loc"], idx, w)
        pred = P.protein_loc_correction(lc, 0.1)
        outs.append((lc.detach().cpu(), loss.item(), pred.cpu(), [p.grad.cpu() for p in m.parameters()]))
    assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] and torch.equal(outs[0][2], outs[1][2])
    for a, b in zip(outs[0][3], outs[1][3]):
        assert rel_err(a, b) < 1e-5                     # the scatter's fp32 reductions are unordered


def test_second_backward_through_the_engine_raises(cuda):
    prob, g, go, m, mo = build_pair(cuda, n=800, e=16000, dims=(3, 20, 20))
    y = m(g, g.ndata["feat"])
    y.sum().backward(retain_graph=True)
    with pytest.raises(P.PlagnnError):
        y.sum().backward()
