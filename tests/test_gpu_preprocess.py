"""Preprocessing kernels against the reference's own outputs (tests/golden/preprocess.npz) and the
oracle on larger seeded inputs.  Integer work and IEEE-exact float64 element operations -> bit-exact comparisons, entry
order included; only the mean / standard deviation of the difference matrix is compared to a tolerance (1e-12 relative:
numpy sums pairwise, the kernel over a fixed two-level tree)."""
import os

import numpy as np
import pytest
import torch
from scipy.sparse import coo_matrix

import plagnn_b200 as P
from oracle import preprocess_oracle as po
from plagnn_b200 import preprocess as pp
from plagnn_b200 import synth
from tests.test_oracle_preprocess import ECC_CASES, MOD_CASES, ppi_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def same_coo(a, b):
    return np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col) and np.array_equal(a.data, b.data)


@pytest.mark.parametrize("name", ECC_CASES)
def test_ecc_bit_exact_vs_reference(cuda, g, name):
    e = pp.edge_clustering_coefficients(ppi_of(g, f"ecc_{name}"), float(g[f"ecc_{name}_eps"]), device=cuda)
    assert np.array_equal(e.row, g[f"ecc_{name}_row"]) and np.array_equal(e.col, g[f"ecc_{name}_col"])
    assert e.data.dtype == np.float64 and np.array_equal(e.data, g[f"ecc_{name}_data"])


@pytest.mark.parametrize("n,m,form", [(3000, 60000, "normal"), (1001, 30000, "inter")])
def test_ecc_bit_exact_vs_oracle(cuda, n, m, form):
    prob = synth.ppi_problem(n, m, form, 11, feat_dims=(3, 4, 4))
    ppi = prob.scipy_ppi()
    assert same_coo(pp.edge_clustering_coefficients(ppi, 0, device=cuda), po.edge_clustering_coefficients(ppi, 0))


def test_ecc_full_size_properties(cuda):
    """PPI-shaped graph (24 041 nodes, 1.4 M directed edges): one entry per stored edge, mirrored pairs equal, values in
    [0, 1] (common neighbours of i and j exclude i and j), and 300 sampled edges recounted with numpy."""
    prob = synth.ppi_problem(24041, 1400000, "normal", 70, feat_dims=(3, 4, 4))
    ppi = prob.scipy_ppi()
    row = torch.from_numpy(ppi.row.astype(np.int32)).to(cuda)
    col = torch.from_numpy(ppi.col.astype(np.int32)).to(cuda)
    r, c, d = (t.cpu().numpy() for t in pp.ecc_device(row, col, 24041, 0.0))
    assert r.size == ppi.nnz
    assert np.array_equal(r[0::2], c[1::2]) and np.array_equal(c[0::2], r[1::2]) and np.array_equal(d[0::2], d[1::2])
    assert np.all(r[0::2] < c[0::2]) and np.all(np.diff(r[0::2].astype(np.int64) * 24041 + c[0::2]) > 0)
    assert d.min() >= 0.0 and d.max() <= 1.0
    csr = po.sorted_csr(ppi)
    deg = np.diff(csr.indptr)
    for k in np.random.default_rng(3).choice(r.size // 2, 300, replace=False):
        i, j = int(r[2 * k]), int(c[2 * k])
        tri = np.intersect1d(csr.indices[csr.indptr[i]:csr.indptr[i + 1]], csr.indices[csr.indptr[j]:csr.indptr[j + 1]]).size
        den = min(deg[i], deg[j]) - 1
        assert d[2 * k] == (0.0 if den == 0 else tri / den)


def test_ecc_duplicate_entry_raises(cuda):
    m = coo_matrix((np.ones(5, dtype=np.int64), ([0, 1, 0, 1, 2], [1, 0, 1, 2, 1])), shape=(4, 4))
    with pytest.raises(P.PlagnnError):
        pp.edge_clustering_coefficients(m, device=cuda)


@pytest.mark.parametrize("name", MOD_CASES)
def test_rewiring_vs_reference(cuda, g, name):
    ppi, thr = ppi_of(g, f"mod_{name}"), float(g[f"mod_{name}_thr"])
    nor, inter = g[f"mod_{name}_pcc_nor"], g[f"mod_{name}_pcc_inter"]
    mean, std = pp.diff_moments(torch.tensor(nor, device=cuda), torch.tensor(inter, device=cuda))
    rm, rs = g[f"mod_{name}_mean_std"]
    assert abs(mean - rm) <= 1e-12 * max(abs(rm), rs) and abs(std - rs) <= 1e-12 * rs
    # the decision kernel with the reference's own thresholds: bit-identical adjacency
    res = pp.modify_network_topology(ppi, nor, inter, thr, device=cuda, thresholds=po.thresholds(rm, rs, thr))
    assert res.data.dtype == np.int64 and np.array_equal(res.data, g[f"mod_{name}_data"])
    assert np.array_equal(res.row, g[f"mod_{name}_row"]) and np.array_equal(res.col, g[f"mod_{name}_col"])
    # the whole call (thresholds from the device moments)
    res = pp.modify_network_topology(ppi, coo_matrix(nor), coo_matrix(inter), thr, device=cuda)
    assert np.array_equal(res.row, g[f"mod_{name}_row"]) and np.array_equal(res.col, g[f"mod_{name}_col"])


@pytest.mark.parametrize("n,m", [(700, 9000), (1024, 20000), (37, 0)])
def test_rewiring_bit_exact_vs_oracle(cuda, n, m):
    rng = np.random.default_rng(n)
    if m:
        ppi = synth.ppi_problem(n, m, "normal", 5, feat_dims=(3, 4, 4)).scipy_ppi()
    else:
        ppi = coo_matrix((n, n), dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mats = [np.nan_to_num(np.corrcoef(rng.normal(8, 2, size=(n, 3)))) for _ in range(2)]
    for p in mats:
        np.fill_diagonal(p, 0)
    want = po.modify_network_topology(ppi, mats[0], mats[1], 1.5)
    lo, hi = po.thresholds(*po.diff_moments(mats[0], mats[1]), 1.5)
    got = pp.modify_network_topology(ppi, mats[0], mats[1], 1.5, device=cuda, thresholds=(lo, hi))
    assert same_coo(got, want) and got.nnz > 0
    mean, std = pp.diff_moments(torch.tensor(mats[0], device=cuda), torch.tensor(mats[1], device=cuda))
    rm, rs = po.diff_moments(mats[0], mats[1])
    assert abs(mean - rm) <= 1e-12 * max(abs(rm), rs) and abs(std - rs) <= 1e-12 * rs


def test_pearson_matrix_against_the_reference_statements(cuda, golden_dir):
    """plagnn_pearson (np.corrcoef tail of construct_gcn_matrix, code/data_preprocess.py:165-170) against the fixture made from
    the reference's own statements: 4 ulp on values in [-1, 1], zero diagonal, zero-variance rows -> 0."""
    import os
    from plagnn_b200 import preprocess
    z = np.load(os.path.join(golden_dir, "pearson.npz"))
    got = preprocess.pearson_matrix(z["expr"], cuda).cpu().numpy()
    assert got.shape == z["pcc"].shape and np.isfinite(got).all() and (np.diag(got) == 0).all()
    assert np.abs(got - z["pcc"]).max() <= 4 * 2.0 ** -53
    zero_var = z["expr"].std(1) == 0
    assert (got[zero_var] == 0).all() and (got[:, zero_var] == 0).all()
    assert np.abs(got - got.T).max() <= 2 * 2.0 ** -53      # (c / sd_i) / sd_j vs (c / sd_j) / sd_i: numpy's is not symmetric to the bit either
    # and it feeds the rewiring step like the reference's matrix does: same thresholds, same rewired network
    rng = np.random.default_rng(5)
    n = got.shape[0]
    expr2 = z["expr"] + rng.normal(0, 0.5, size=z["expr"].shape) * (z["expr"].std(1, keepdims=True) > 0)
    pcc2_ref = __import__("oracle.preprocess_oracle", fromlist=["x"]).pearson_matrix(expr2)
    pcc2 = preprocess.pearson_matrix(expr2, cuda)
    m_ref = (float(np.mean(pcc2_ref - z["pcc"])), float(np.std(pcc2_ref - z["pcc"])))
    m_got = preprocess.diff_moments(torch.from_numpy(got).to(cuda), pcc2)
    assert abs(m_got[0] - m_ref[0]) < 1e-12 and abs(m_got[1] - m_ref[1]) < 1e-12


def test_pearson_full_ppi_shape(cuda):
    """N = 24 041 proteins x 3 samples (BASELINE.json configs[0-1] shape): 4.6 GB matrix, spot rows against numpy."""
    from plagnn_b200 import preprocess
    rng = np.random.default_rng(11)
    n = 24041
    expr = np.abs(rng.normal(8.0, 2.0, size=(n, 3)))
    expr[rng.random(n) < 0.1] = 0.0
    out = preprocess.pearson_matrix(expr, cuda)
    rows = [0, 1, 777, 12000, n - 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        for r in rows:
            x = expr - expr.mean(1, keepdims=True)
            c = (x @ x[r]) * 0.5
            sd = np.sqrt((x * x).sum(1) * 0.5)
            ref = np.clip(c / sd[r] / sd, -1, 1)
            ref[r] = 0
            ref[np.isnan(ref)] = 0
            assert np.abs(out[r].cpu().numpy() - ref).max() <= 8 * 2.0 ** -53
