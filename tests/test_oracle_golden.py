"""Pins the oracle against vectors produced by the REFERENCE's own functions
(tests/golden/make_golden.py ran code/train.py and code/main.py from /root/reference)."""
import os

import numpy as np
import torch

from oracle import plagnn_oracle as orc


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_functions.npz"))


def test_multi_loss_matches_reference_value_and_grad(golden_dir):
    g = _g(golden_dir)
    p = torch.tensor(g["loss_p"], requires_grad=True)
    loss = orc.multi_loss(p, torch.tensor(g["loss_t"]), g["loss_w"])
    loss.backward()
    assert loss.dtype == torch.float32
    # same operator order as the reference -> bit-identical on the same torch build
    assert np.array_equal(loss.detach().numpy(), g["loss_value"])
    assert np.array_equal(p.grad.numpy(), g["loss_grad"])


def test_multi_loss_saturation_cases(golden_dir):
    g = _g(golden_dir)
    # p == 0 -> clamp kills the gradient of the positive branch; p == 1 -> the negative one
    assert g["loss_p"][0, 0] == 0.0 and g["loss_p"][0, 1] == 1.0
    grad = g["loss_grad"]
    t = g["loss_t"]
    if t[0, 0] == 1:
        assert grad[0, 0] == 0.0
    assert np.isfinite(grad).all()


def test_weight_cal(golden_dir):
    g = _g(golden_dir)
    assert np.array_equal(orc.weight_cal(g["wc_loc"]), g["wc_out"])


def test_label_decision_and_metrics(golden_dir):
    g = _g(golden_dir)
    pred = orc.protein_loc_correction(torch.tensor(g["lc_probs"]), float(g["lc_alpha"]))
    assert pred.dtype == torch.float64
    assert np.array_equal(pred.numpy(), g["lc_pred"])
    aim, cov, acc = orc.performances_record(torch.tensor(g["pr_truth"]), torch.tensor(g["lc_pred"]))
    np.testing.assert_allclose([aim, cov, acc], g["pr_out"], rtol=2e-6)


def test_scaling(golden_dir):
    g = _g(golden_dir)
    np.testing.assert_allclose(orc.scaling(g["sc_in"]), g["sc_out"], rtol=1e-15)


# ---- alteration scoring (tests/golden/make_golden_scoring.py ran code/main.py's own statements) ----
def _s(golden_dir):
    return np.load(os.path.join(golden_dir, "scoring.npz"))


def test_scaling_float32_and_float64_bit_exact(golden_dir):
    g = _s(golden_dir)
    out32 = orc.scaling(g["sc32_in"])
    assert out32.dtype == np.float32 and np.array_equal(out32, g["sc32_out"])
    assert np.array_equal(orc.scaling(g["sc64_in"]), g["sc64_out"])


def test_mat_merge_bit_exact(golden_dir):
    g = _s(golden_dir)
    assert np.array_equal(orc.mat_merge(list(g["runs"])), g["merged"])


def check_ranking(diff_flat, order, ref_order):
    """Same score at every rank as the reference (NaN positions included), a valid permutation, and the same index
    wherever the score is unique."""
    assert sorted(order.tolist()) == list(range(len(diff_flat)))
    a, b = diff_flat[order], diff_flat[ref_order]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    vals, counts = np.unique(diff_flat[~np.isnan(diff_flat)], return_counts=True)
    unique_vals = set(vals[counts == 1].tolist())
    same = [i for i in range(len(order)) if diff_flat[ref_order[i]] in unique_vals]
    assert np.array_equal(order[same], ref_order[same])


def test_alteration_rank_matches_reference(golden_dir):
    g = _s(golden_dir)
    normal, inter, diff, order = orc.alteration_rank(g["normal_mat"], g["inter_mat"])
    assert np.array_equal(normal, g["normal"]) and np.array_equal(inter, g["inter"])
    assert np.array_equal(diff, g["diff"], equal_nan=True)
    check_ranking(diff.reshape(-1), order, g["order"])
