"""Alteration scoring on the device against the reference's own outputs (tests/golden/scoring.npz)
and against the oracle on larger seeded inputs.  Element arithmetic is IEEE-exact -> bit-exact comparisons."""
import os

import numpy as np
import pytest
import torch

import plagnn_b200 as P
from oracle import plagnn_oracle as orc
from plagnn_b200 import scoring
from tests.test_oracle_golden import check_ranking

pytestmark = pytest.mark.gpu


def _s(golden_dir):
    return np.load(os.path.join(golden_dir, "scoring.npz"))


def test_scaling_bit_exact_vs_reference(cuda, golden_dir):
    g = _s(golden_dir)
    out32 = scoring.scaling(torch.tensor(g["sc32_in"], device=cuda))
    assert out32.dtype == torch.float32 and np.array_equal(out32.cpu().numpy(), g["sc32_out"])
    out64 = scoring.scaling(torch.tensor(g["sc64_in"], device=cuda))
    assert out64.dtype == torch.float64 and np.array_equal(out64.cpu().numpy(), g["sc64_out"])


def test_mat_merge_bit_exact_vs_reference(cuda, golden_dir):
    g = _s(golden_dir)
    runs = [torch.tensor(m, device=cuda) for m in g["runs"]]
    assert np.array_equal(scoring.mat_merge(runs).cpu().numpy(), g["merged"])


def test_alteration_rank_vs_reference(cuda, golden_dir):
    g = _s(golden_dir)
    normal, inter, diff, order = scoring.alteration_rank(torch.tensor(g["normal_mat"], device=cuda),
                                                         torch.tensor(g["inter_mat"], device=cuda))
    assert np.array_equal(normal.cpu().numpy(), g["normal"]) and np.array_equal(inter.cpu().numpy(), g["inter"])
    assert np.array_equal(diff.cpu().numpy(), g["diff"], equal_nan=True)
    check_ranking(g["diff"].reshape(-1), order.cpu().numpy(), g["order"])


@pytest.mark.parametrize("rows,cols", [(24041, 12), (1000, 7), (513, 20), (40, 128)])
def test_scoring_full_size_vs_oracle(cuda, rows, cols):
    rng = np.random.default_rng(rows + cols)
    normal_mat = rng.uniform(0.01, 0.99, size=(rows, cols))
    inter_mat = normal_mat * rng.uniform(0.7, 1.4, size=(rows, cols))
    if rows > 10:
        inter_mat[5] = normal_mat[5]
        inter_mat[7, :] = inter_mat[6, :] = 0.5                         # exact ties
        normal_mat[7, :] = normal_mat[6, :] = 0.25
    n_o, i_o, d_o, o_o = orc.alteration_rank(normal_mat, inter_mat)
    normal, inter, diff, order = scoring.alteration_rank(torch.tensor(normal_mat, device=cuda), torch.tensor(inter_mat, device=cuda))
    assert np.array_equal(normal.cpu().numpy(), n_o) and np.array_equal(inter.cpu().numpy(), i_o)
    assert np.array_equal(diff.cpu().numpy(), d_o, equal_nan=True)
    assert np.array_equal(order.cpu().numpy(), o_o)                      # same tie rule as the oracle: fully identical
    # size-independent properties: sortedness and permutation
    flat = diff.reshape(-1)[order].cpu().numpy()
    finite = flat[~np.isnan(flat)]
    assert (finite[1:] <= finite[:-1]).all()
    assert np.isnan(flat[: int(np.isnan(flat).sum())]).all()


def test_mat_merge_full_size_vs_oracle_and_records(cuda):
    rng = np.random.default_rng(9)
    runs = rng.uniform(0.001, 0.999, size=(10, 24041, 12)).astype(np.float32)
    got = scoring.mat_merge([torch.tensor(m, device=cuda) for m in runs])
    assert np.array_equal(got.cpu().numpy(), orc.mat_merge(list(runs)))
    inter = got * torch.tensor(rng.uniform(0.8, 1.2, size=(24041, 12)), device=cuda)
    rec = scoring.misloc_records(got, inter, top=1000)
    assert rec["rank"][0].item() == 1 and rec["row"].numel() == 1000
    s = rec["score"].cpu().numpy()
    assert (s != -1.0).all() and (s != 0).all() and not np.isnan(s).any()
    assert (s[1:] <= s[:-1]).all()


def test_scoring_rejects_cpu_tensors():
    with pytest.raises(P.PlagnnError):
        scoring.scaling(torch.zeros(4, 12))
