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


def ppi_of(g, prefix):
    n = int(g[f"{prefix}_n"])
    r, c = g[f"{prefix}_in_row"], g[f"{prefix}_in_col"]
    return coo_matrix((np.ones(r.size, dtype=np.int64), (r, c)), shape=(n, n))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


@pytest.mark.parametrize("name", ECC_CASES)
def test_ecc_oracle_equals_reference(g, name):
    e = po.edge_clustering_coefficients(ppi_of(g, f"ecc_{name}"), float(g[f"ecc_{name}_eps"]))
    assert np.array_equal(e.row, g[f"ecc_{name}_row"]) and np.array_equal(e.col, g[f"ecc_{name}_col"])
    assert e.data.dtype == np.float64 and np.array_equal(e.data, g[f"ecc_{name}_data"])


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
