"""oracle/preprocess_oracle.py against the outputs of the reference's own `edge_clustering_coefficients` and
`modify_network_topology` (tests/golden/preprocess.npz, made by tests/golden/make_golden_preprocess.py from
code/data_preprocess.py:175-257): bit-exact, entry order included."""
import os

import numpy as np
import pytest
from scipy.sparse import coo_matrix

from oracle import preprocess_oracle as po

ECC_CASES = ("powerlaw", "powerlaw_eps", "star", "k6", "path_isolated", "empty")
MOD_CASES = ("a", "b", "c")
# fixtures added after the last GPU run of round 1: checked against the oracle here, to be added to the GPU list (ECC_CASES,
# which tests/test_gpu_preprocess.py imports) with the next GPU run
ECC_CASES_ORACLE_ONLY = ("diagonal", "asymmetric")


def ppi_of(g, prefix):
    n = int(g[f"{prefix}_n"])
    r, c = g[f"{prefix}_in_row"], g[f"{prefix}_in_col"]
    return coo_matrix((np.ones(r.size, dtype=np.int64), (r, c)), shape=(n, n))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


@pytest.mark.parametrize("name", ECC_CASES + ECC_CASES_ORACLE_ONLY)
def test_ecc_oracle_equals_reference(g, name):
    e = po.edge_clustering_coefficients(ppi_of(g, f"ecc_{name}"), float(g[f"ecc_{name}_eps"]))
    assert np.array_equal(e.row, g[f"ecc_{name}_row"]) and np.array_equal(e.col, g[f"ecc_{name}_col"])
    assert e.data.dtype == np.float64 and np.array_equal(e.data, g[f"ecc_{name}_data"])
    assert np.array_equal(np.signbit(e.data), np.signbit(g[f"ecc_{name}_data"]))      # 0 / -1 = -0.0 in the asymmetric case


def test_ecc_known_answers(g):
    # K6: every edge closes 4 triangles, min degree 5 -> 4 / 4; star: denominators are 0 -> epsilon
    assert np.all(g["ecc_k6_data"] == 1.0) and g["ecc_k6_data"].size == 30
    assert np.all(g["ecc_star_data"] == 0.0) and g["ecc_star_data"].size == 22
    assert np.all(g["ecc_path_isolated_data"][[0, 1, 4, 5]] == 7.5)      # (1,2) and (5,6): a leaf on one side
    assert np.all(g["ecc_path_isolated_data"][[2, 3]] == 7.5)            # (2,3): min(2, 1) - 1 = 0 as well


@pytest.mark.parametrize("name", MOD_CASES)
def test_rewiring_oracle_equals_reference(g, name):
    res = po.modify_network_topology(ppi_of(g, f"mod_{name}"), g[f"mod_{name}_pcc_nor"], g[f"mod_{name}_pcc_inter"],
                                     float(g[f"mod_{name}_thr"]))
    assert np.array_equal(res.row, g[f"mod_{name}_row"]) and np.array_equal(res.col, g[f"mod_{name}_col"])
    assert res.data.dtype == np.int64 and np.array_equal(res.data, g[f"mod_{name}_data"])
    assert np.array_equal(np.array(po.diff_moments(g[f"mod_{name}_pcc_nor"], g[f"mod_{name}_pcc_inter"])), g[f"mod_{name}_mean_std"])


def brute_force_ecc(adj, epsilon):
    """the definition on a dense 0/1 matrix, one edge at a time (small n only)"""
    n = adj.shape[0]
    deg = adj.sum(axis=1)
    rows, cols, vals = [], [], []
    for i in range(n):
        for j in range(i + 1, n):
            if adj[i, j]:
                den = min(deg[i], deg[j]) - 1
                v = float(epsilon) if den == 0 else np.float64(np.count_nonzero(adj[i] & adj[j])) / np.float64(den)
                rows += [i, j]
                cols += [j, i]
                vals += [v, v]
    return np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32), np.array(vals, dtype=np.float64)


@pytest.mark.parametrize("seed", range(6))
def test_ecc_oracle_equals_dense_definition(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 40))
    adj = rng.random((n, n)) < rng.uniform(0.05, 0.6)
    adj = np.triu(adj, 1)
    adj = (adj | adj.T).astype(np.int64)
    if seed % 2:
        adj[np.arange(0, n, 3), np.arange(0, n, 3)] = 1          # some self-interactions
    if seed >= 3:
        adj[rng.random((n, n)) < 0.05] = 0                        # a few one-directional entries (rewired networks)
    r, c = np.nonzero(adj)
    perm = rng.permutation(r.size)
    m = coo_matrix((np.ones(r.size, dtype=np.int64), (r[perm].astype(np.int32), c[perm].astype(np.int32))), shape=(n, n))
    e = po.edge_clustering_coefficients(m, 0.5)
    br, bc, bv = brute_force_ecc(adj, 0.5)
    assert np.array_equal(e.row, br) and np.array_equal(e.col, bc) and np.array_equal(e.data, bv)


def test_rewiring_thresholds_are_strict(g):
    """a difference equal to a threshold changes nothing (`<` and `>` at code/data_preprocess.py:250-251)"""
    n = 4
    ppi = coo_matrix((np.ones(2, dtype=np.int64), ([0, 1], [1, 0])), shape=(n, n))
    nor = np.zeros((n, n))
    inter = np.zeros((n, n))
    inter[0, 1] = inter[1, 0] = -0.5          # existing edge at the left threshold: kept
    inter[2, 3] = inter[3, 2] = 0.5           # missing edge at the right threshold: not added
    inter[0, 2] = 0.75                        # missing, above: added (one direction only, as the reference would)
    res = po.modify_network_topology(ppi, nor, inter, 0.0, l_threshold=-0.5, r_threshold=0.5)
    assert sorted(zip(res.row.tolist(), res.col.tolist())) == [(0, 1), (0, 2), (1, 0)]


def test_rewiring_changes_both_ways(g):
    """the fixtures exercise removal and insertion (a vector that only added edges would not pin the first rule)"""
    for name in MOD_CASES:
        n = int(g[f"mod_{name}_n"])
        before = np.zeros((n, n), dtype=bool)
        before[g[f"mod_{name}_in_row"], g[f"mod_{name}_in_col"]] = True
        after = np.zeros((n, n), dtype=bool)
        after[g[f"mod_{name}_row"], g[f"mod_{name}_col"]] = True
        assert (before & ~after).any() and (~before & after).any(), name


def test_device_mirror_refuses_cpu():
    import plagnn_b200 as P
    m = coo_matrix((np.ones(2, dtype=np.int64), ([0, 1], [1, 0])), shape=(3, 3))
    with pytest.raises(P.PlagnnError):
        P.preprocess.edge_clustering_coefficients(m, device="cpu")
    with pytest.raises(P.PlagnnError):
        P.preprocess.modify_network_topology(m, np.zeros((3, 3)), np.zeros((3, 3)), 1.0, device="cpu")


def test_workspace_queries():
    from plagnn_b200 import _lib
    lib = _lib.load()
    assert lib.plagnn_ecc_workspace_bytes(24041) >= 3 * 4 * 24042
    assert lib.plagnn_ecc_workspace_bytes(0) == 0
    assert lib.plagnn_rewire_workspace_bytes(24041) >= 4 * 24041


def test_pearson_oracle_against_the_reference_statements(golden_dir):
    """np.corrcoef tail of construct_gcn_matrix (code/data_preprocess.py:165-170), fixture from the reference's own statements."""
    import os
    z = np.load(os.path.join(golden_dir, "pearson.npz"))
    got = po.pearson_matrix(z["expr"])
    assert got.shape == z["pcc"].shape and np.isfinite(got).all() and (np.diag(got) == 0).all()
    assert np.abs(got - z["pcc"]).max() <= 4 * 2.0 ** -53          # 4 ulp at 1.0 (BLAS builds may differ in the last bits)
    zero_var = z["expr"].std(1) == 0
    assert zero_var.sum() > 5 and (got[zero_var] == 0).all() and (got[:, zero_var] == 0).all()
