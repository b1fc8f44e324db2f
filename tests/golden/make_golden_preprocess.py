"""Golden vectors for the preprocessing step, produced by the REFERENCE's own functions: the module
code/data_preprocess.py is imported unchanged (its pipeline sits behind `if __name__ == '__main__'`) and
`edge_clustering_coefficients` (:175-214) and `modify_network_topology` (:217-257) are run on small seeded inputs.
Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_preprocess.py        -> tests/golden/preprocess.npz
"""
import os
import sys

import numpy as np
from scipy.sparse import coo_matrix

sys.path.insert(0, "/root/reference/code")
import data_preprocess as ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(4203)


def sym_graph(n, pairs):
    """symmetric 0/1 COO with int64 ones, entries in shuffled order (PPI_normal is built from a Python set,
    code/data_preprocess.py:83-108)"""
    pairs = sorted({(min(a, b), max(a, b)) for a, b in pairs if a != b})
    r = np.array([p[0] for p in pairs] + [p[1] for p in pairs], dtype=np.int32)
    c = np.array([p[1] for p in pairs] + [p[0] for p in pairs], dtype=np.int32)
    perm = rng.permutation(r.size)
    return coo_matrix((np.ones(r.size, dtype=np.int64), (r[perm], c[perm])), shape=(n, n))


def powerlaw_pairs(n, m):
    w = 1.0 / np.arange(1, n + 1) ** 0.8
    w /= w.sum()
    a = rng.choice(n, size=m, p=w)
    b = rng.choice(n, size=m, p=w)
    return list(zip(a.tolist(), b.tolist()))


cases = {
    "powerlaw": (sym_graph(90, powerlaw_pairs(90, 700)), 0),
    "powerlaw_eps": (sym_graph(64, powerlaw_pairs(64, 260)), 0.25),
    "star": (sym_graph(12, [(0, k) for k in range(1, 12)]), 0),                       # no triangles, leaves of degree 1
    "k6": (sym_graph(6, [(a, b) for a in range(6) for b in range(a + 1, 6)]), 0),    # every edge in 4 triangles
    "path_isolated": (sym_graph(9, [(1, 2), (2, 3), (5, 6)]), 7.5),                   # denominators 0 -> epsilon; isolated nodes
    "empty": (coo_matrix((5, 5), dtype=np.int64), 0),
}


def with_diagonal(g, nodes):
    """self-interactions kept on the diagonal: not edges of the loop (`neighbors > i`, :195) but counted in the degree and
    in the row overlap"""
    r = np.concatenate([g.row, np.array(nodes, dtype=np.int32)])
    c = np.concatenate([g.col, np.array(nodes, dtype=np.int32)])
    return coo_matrix((np.ones(r.size, dtype=np.int64), (r, c)), shape=g.shape)

out = {}


def record_ecc(name, g, eps):
    ecc = ref.edge_clustering_coefficients(g, epsilon=eps)
    out[f"ecc_{name}_in_row"], out[f"ecc_{name}_in_col"] = g.row.astype(np.int32), g.col.astype(np.int32)
    out[f"ecc_{name}_n"], out[f"ecc_{name}_eps"] = np.int64(g.shape[0]), np.float64(eps)
    out[f"ecc_{name}_row"], out[f"ecc_{name}_col"] = ecc.row.astype(np.int32), ecc.col.astype(np.int32)
    out[f"ecc_{name}_data"] = ecc.data.astype(np.float64)


for name, (g, eps) in cases.items():
    record_ecc(name, g, eps)
    # (when every value is the integer epsilon, e.g. the star, scipy infers int64 for the list; values are what is pinned)


def pcc(n, samples=3):
    """what construct_gcn_matrix leaves (code/data_preprocess.py:166-172): Pearson matrix, zero diagonal, NaN -> 0"""
    x = rng.normal(8.0, 2.0, size=(n, samples))
    x[rng.random(n) < 0.1] = 0.0                       # proteins without expression: constant rows -> NaN -> 0
    with np.errstate(invalid="ignore", divide="ignore"):
        p = np.corrcoef(x)
    np.fill_diagonal(p, 0)
    p[np.isnan(p)] = 0
    return coo_matrix(p)


for name, n, m, thr in (("a", 70, 420, 1.0), ("b", 45, 150, 2.0), ("c", 33, 60, 0.5)):
    g = sym_graph(n, powerlaw_pairs(n, m))
    pn, pi = pcc(n), pcc(n)
    res = ref.modify_network_topology(g, pn, pi, thr)
    diff = (pi.tocsr() - pn.tocsr()).toarray()
    out[f"mod_{name}_in_row"], out[f"mod_{name}_in_col"] = g.row.astype(np.int32), g.col.astype(np.int32)
    out[f"mod_{name}_n"], out[f"mod_{name}_thr"] = np.int64(n), np.float64(thr)
    out[f"mod_{name}_pcc_nor"], out[f"mod_{name}_pcc_inter"] = pn.toarray(), pi.toarray()
    out[f"mod_{name}_mean_std"] = np.array([np.mean(diff), np.std(diff)])
    out[f"mod_{name}_row"], out[f"mod_{name}_col"] = res.row.astype(np.int32), res.col.astype(np.int32)
    out[f"mod_{name}_data"] = res.data
    print(name, "edges", g.nnz, "->", res.nnz, "dtype", res.data.dtype)

# drawn last so that the generator state of the cases above does not depend on it
record_ecc("diagonal", with_diagonal(sym_graph(40, powerlaw_pairs(40, 160)), [0, 1, 2, 5, 17, 39]), 0)

# the rewired network can be asymmetric in a few entries (ECC is applied to it at code/data_preprocess.py:327): entries
# (i, j) without (j, i), one of them pointing at a node whose own row is empty (denominator -1 -> value -0.0)
g_asym = sym_graph(30, powerlaw_pairs(30, 90))
keep = ~(((g_asym.row == 7) & (g_asym.col < 7)) | ((g_asym.row == 12) & (g_asym.col < 12)))
extra_r, extra_c = np.array([3, 4], dtype=np.int32), np.array([29, 29], dtype=np.int32)
sel = keep & (g_asym.row != 29) & (g_asym.col != 29)
g_asym = coo_matrix((np.ones(sel.sum() + 2, dtype=np.int64), (np.concatenate([g_asym.row[sel], extra_r]),
                                                              np.concatenate([g_asym.col[sel], extra_c]))), shape=(30, 30))
record_ecc("asymmetric", g_asym, 0)

np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **out)
print("wrote", os.path.join(OUT, "preprocess.npz"), os.path.getsize(os.path.join(OUT, "preprocess.npz")), "bytes")
