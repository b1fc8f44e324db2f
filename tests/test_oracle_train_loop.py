"""Re-runs the epoch loop with the oracle's own restatement (train_epoch) and compares with the fixture
written by the reference's UNCHANGED train.train() (tests/golden/make_golden.py, part 5)."""
import os
import random

import numpy as np
import torch
from scipy.sparse import coo_matrix
from sklearn.model_selection import KFold

from oracle import plagnn_oracle as orc


def load_problem(golden_dir):
    z = np.load(os.path.join(golden_dir, "train_loop_tiny.npz"))
    n = z["loc"].shape[0]
    ppi = coo_matrix((np.ones(len(z["ppi_row"]), dtype=np.int64), (z["ppi_row"], z["ppi_col"])), shape=(n, n))
    return z, ppi, n


def run_loop(z, make_graph, make_model, make_opt, device, epoch_fn, multi_loss):
    """train.py:162-207 restated for round 1 (fold seed 12): returns per-fold loss lists and final logits."""
    seed = int(z["seed"])
    random.seed(seed); torch.manual_seed(seed); np.random.seed(seed)
    g = make_graph()
    features, labels = g.ndata["feat"], g.ndata["loc"]
    i_weight = orc.weight_cal(z["loc"])
    label = [int(i) for i in z["labelled"]]
    kfold = KFold(n_splits=int(z["folds"]), random_state=int(z["fold_seed"]), shuffle=True)
    tl, vl, lg = [], [], []
    for train_idx, val_idx in kfold.split(label):
        model = make_model(features.shape[1]).to(device)
        opt = make_opt(model.parameters(), float(z["lr"]))
        train_index = [label[i] for i in train_idx]
        val_index = [label[i] for i in val_idx]
        t_l, v_l = [], []
        for _ in range(int(z["epochs"])):
            logits, loss = epoch_fn(model, opt, g, features, labels, train_index, i_weight)
            with torch.no_grad():
                v = multi_loss(logits[val_index], labels[val_index], i_weight)
            t_l.append(float(loss))
            v_l.append(float(v))
        tl.append(t_l); vl.append(v_l); lg.append(logits.detach().float().cpu().numpy())
    return np.array(tl), np.array(vl), np.stack(lg)


def test_oracle_epoch_loop_reproduces_reference_run(golden_dir):
    z, ppi, n = load_problem(golden_dir)
    uniprot = list(range(n))
    tl, vl, lg = run_loop(
        z, lambda: orc.create_graph(ppi, z["ecc"], z["gcn"], coo_matrix(z["loc"]), z["expr"], uniprot),
        lambda f: orc.GNN32Ref(f, 400, 300, 200, 100, 12), lambda p, lr: torch.optim.Adam(p, lr=lr), "cpu",
        orc.train_epoch, orc.multi_loss)
    # identical code path (oracle under the reference loop) -> identical numbers on the same torch build
    np.testing.assert_allclose(tl, z["train_loss"], rtol=1e-6)
    np.testing.assert_allclose(vl, z["val_loss"], rtol=1e-6)
    np.testing.assert_allclose(lg, z["logits"], rtol=1e-5, atol=1e-7)


def test_golden_csc_matches_oracle(golden_dir):
    z, ppi, n = load_problem(golden_dir)
    s, d = orc.add_self_loop(ppi.row.astype(np.int64), ppi.col.astype(np.int64), n)
    indptr, indices, eids = orc.coo_to_csc(s, d, n)
    assert np.array_equal(indptr, z["indptr"]) and np.array_equal(indices, z["indices"]) and np.array_equal(eids, z["eids"])
