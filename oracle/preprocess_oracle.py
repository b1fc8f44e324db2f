"""CPU restatement of the two preprocessing steps the device kernels replace.  TEST INFRASTRUCTURE
ONLY: imported by tests/ (and nothing in the product path); vectorised numpy/scipy, written from the reference's
semantics, not its loops.  Pinned: tests/golden/preprocess.npz holds outputs of the reference's own functions
(tests/golden/make_golden_preprocess.py imports code/data_preprocess.py unchanged); tests/test_oracle_golden.py checks
this file against them bit for bit.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse import coo_matrix


def sorted_csr(ppi_net):
    """`ppi_net.tocsr()` (code/data_preprocess.py:184): duplicates summed, columns ascending inside a row."""
    m = ppi_net.tocsr().copy()
    m.sum_duplicates()
    m.sort_indices()
    return m


def edge_clustering_coefficients(ppi_net, epsilon=0):
    """code/data_preprocess.py:175-214.  For each stored (i, j) with j > i, rows ascending and columns ascending inside a
    row (:190,:195): triangles = number of columns where both rows are non-zero (:197), degree = sum of the row's data
    (:192,:198), value = triangles / (min(deg_i, deg_j) - 1), epsilon when that denominator is 0 (:199-203); the lists
    receive (i, j, v) then (j, i, v) (:205-210).  Returned data is float64 (the reference's list gives int64 only when
    every value is an integer epsilon)."""
    ppi = sorted_csr(ppi_net)
    n = ppi.shape[0]
    pattern = sparse.csr_matrix((np.ones(ppi.nnz, dtype=np.int64) * (ppi.data != 0), ppi.indices, ppi.indptr), shape=ppi.shape)
    degree = np.asarray(ppi.sum(axis=1)).ravel().astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ppi.indptr))
    cols = ppi.indices.astype(np.int64)
    up = cols > rows                                    # CSR order is already (i ascending, j ascending)
    i, j = rows[up], cols[up]
    if i.size == 0:
        return coo_matrix((np.zeros(0), (np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))), shape=ppi.shape)
    common = (pattern @ pattern.T).tocsr()
    tri = np.asarray(common[i, j]).ravel().astype(np.int64)
    possible = np.minimum(degree[i], degree[j]) - 1
    value = np.full(i.size, float(epsilon), dtype=np.float64)
    ok = possible != 0
    value[ok] = tri[ok] / possible[ok]                  # int64 / int64 -> float64, correctly rounded
    r = np.empty(2 * i.size, dtype=np.int32)
    c = np.empty(2 * i.size, dtype=np.int32)
    r[0::2], r[1::2] = i, j
    c[0::2], c[1::2] = j, i
    return coo_matrix((np.repeat(value, 2), (r, c)), shape=ppi.shape)


def diff_moments(pcc_nor, pcc_inter):
    """code/data_preprocess.py:236,242-244: mean and (population) standard deviation of the dense difference."""
    diff = _dense(pcc_inter) - _dense(pcc_nor)
    return float(np.mean(diff)), float(np.std(diff))


def thresholds(mean, std, thr):
    """code/data_preprocess.py:245-246."""
    return mean - thr * std, mean + thr * std


def _dense(m):
    return m.toarray() if sparse.issparse(m) else np.asarray(m)


def modify_network_topology(ppi_net, pcc_nor, pcc_inter, thr, l_threshold=None, r_threshold=None):
    """code/data_preprocess.py:217-257.  Elementwise `inter - normal` equals the reference's sparse subtraction (a missing
    entry is an exact 0).  An entry equal to 1 whose difference is below the left threshold becomes 0, an entry equal to 0
    whose difference is above the right threshold becomes 1 (:250-253); result = coo_matrix(dense) (:255): row-major,
    int64 data."""
    adj = np.asarray(ppi_net.tocsr().todense()).astype(np.int64)
    diff = _dense(pcc_inter) - _dense(pcc_nor)
    if l_threshold is None:
        l_threshold, r_threshold = thresholds(np.mean(diff), np.std(diff), thr)
    out = adj.copy()
    out[(diff < l_threshold) & (adj == 1)] = 0
    out[(diff > r_threshold) & (adj == 0)] = 1
    return coo_matrix(out)


def pearson_matrix(expr_gcn: np.ndarray) -> np.ndarray:
    """data_preprocess.py:165-170 (the numeric tail of construct_gcn_matrix): np.corrcoef of the protein x sample matrix,
    diagonal and NaN entries (zero-variance rows) set to 0.  np.corrcoef is the reference's own call; its BLAS product is not
    bit-stable across builds, so comparisons against it carry a 4-ulp bar (values in [-1, 1])."""
    with np.errstate(invalid="ignore", divide="ignore"):
        p = np.corrcoef(np.asarray(expr_gcn, dtype=np.float64))
    np.fill_diagonal(p, 0)
    p[np.isnan(p)] = 0
    return p
