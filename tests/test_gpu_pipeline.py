"""BASELINE.json configs[1] as one pipeline on the device (train control + perturbation state, merge, score, rank:
plagnn_b200.pipeline) against the same pipeline built from the oracle's pieces on the CPU."""
import numpy as np
import pytest
import torch

import plagnn_b200 as P
from plagnn_b200 import pipeline, scoring, synth
from oracle import plagnn_oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

SEEDS, FOLDS, EPOCHS, LR = (12, 22), 2, 3, 5e-5


def _oracle_state(prob, model_seed):
    ids = list(range(prob.num_nodes))
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    w = orc.weight_cal(prob.loc)
    torch.manual_seed(model_seed)
    outs = []
    for seed in SEEDS:
        for train_index, _ in pipeline.fold_splits(prob.labelled, FOLDS, seed):
            mo = orc.GNN32Ref(go.ndata["feat"].shape[1], 400, 300, 200, 100, 12)
            oo = torch.optim.Adam(mo.parameters(), lr=LR)
            idx = [int(i) for i in train_index]
            for _ in range(EPOCHS):
                logits, _ = orc.train_epoch(mo, oo, go, go.ndata["feat"], go.ndata["loc"], idx, w)
            outs.append(logits.detach().numpy())
    return outs


def test_alteration_pipeline_matches_the_oracle_pipeline(cuda):
    n, e, dims = 800, 16000, (3, 40, 40)
    pn = synth.ppi_problem(n, e, "normal", 70, feat_dims=dims)
    pi = synth.ppi_problem(n, e, "inter", 70, feat_dims=dims)
    ids = list(range(n))
    gn = P.create_graph(pn.scipy_ppi(), pn.ecc, pn.gcn, pn.scipy_loc(), pn.expr, ids).to(cuda)
    gi = P.create_graph(pi.scipy_ppi(), pi.ecc, pi.gcn, pi.scipy_loc(), pi.expr, ids).to(cuda)
    rec, normal_m, inter_m = pipeline.alteration_pipeline(gn, gi, pn.labelled, pn.loc, lr=LR, fold_num=FOLDS, epoch_num=EPOCHS,
                                                          fold_seeds=SEEDS, model_seed=70)
    on, oi = _oracle_state(pn, 70), _oracle_state(pi, 70)
    ref_n, ref_i = orc.mat_merge(on), orc.mat_merge(oi)
    assert rel_err(normal_m, ref_n) < 1e-4 and rel_err(inter_m, ref_i) < 1e-4      # three epochs of two fp32 pipelines
    # the scoring stage alone is exact: the oracle's merged matrices through the device kernels give the oracle's ranking
    _, _, diff_ref, order_ref = orc.alteration_rank(ref_n, ref_i)
    _, _, diff_dev, order_dev = scoring.alteration_rank(torch.from_numpy(ref_n).to(cuda), torch.from_numpy(ref_i).to(cuda))
    assert np.array_equal(order_dev.cpu().numpy(), order_ref)
    assert np.array_equal(diff_dev.cpu().numpy(), diff_ref, equal_nan=True)
    # end to end: the ranked table of the device pipeline against the oracle pipeline's
    cols = diff_ref.shape[1]
    flat = diff_ref.reshape(-1)
    keep = [int(i) for i in order_ref if flat[i] != -1.0 and (flat[i] > 0 or flat[i] < 0)]
    top_ref = keep[:50]
    got = (rec["row"] * cols + rec["col"]).cpu().numpy().tolist()
    assert len(got) == len(keep)
    assert len(set(got[:50]) & set(top_ref)) >= 45          # near-equal scores may swap places between fp32 pipelines
    score = dict(zip(got, rec["score"].cpu().numpy().tolist()))
    finite_ref = [i for i in keep if np.isfinite(flat[i])][:50]       # (entries with normal == 0 score +-inf and rank first)
    finite = [i for i in finite_ref if i in score and np.isfinite(score[i])]
    assert len(finite) >= 45
    for i in finite:
        assert abs(score[i] - flat[i]) <= 2e-2 * abs(flat[i]) + 1e-6
    assert rec["rank"][0].item() == 1 and rec["rank"][-1].item() == len(got)
