"""Times the REFERENCE's own `edge_clustering_coefficients` / `modify_network_topology` (code/data_preprocess.py, imported
unchanged from /root/reference — build container only) on the synthetic PPI-shaped inputs that tools/preprocess_time.py
gives the device kernels, and checks on the way that the oracle restatement returns the same matrices.
    python tests/checks/preprocess_cpu_reference.py [--nodes N --edges E --dense-nodes M]"""
import argparse
import json
import os
import sys
import time

import numpy as np
from scipy.sparse import coo_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")
import data_preprocess as ref  # noqa: E402
from oracle import preprocess_oracle as po  # noqa: E402
from plagnn_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=24041)
ap.add_argument("--edges", type=int, default=1400000)
ap.add_argument("--dense-nodes", type=int, default=8192, help="size of the rewiring problem (five dense N x N temporaries)")
a = ap.parse_args()
out = {"cores": os.cpu_count()}

ppi = synth.ppi_problem(a.nodes, a.edges, "normal", 70, feat_dims=(3, 4, 4)).scipy_ppi()
t = time.perf_counter()
e_ref = ref.edge_clustering_coefficients(ppi)
out["ecc_reference_s"] = time.perf_counter() - t
t = time.perf_counter()
e_or = po.edge_clustering_coefficients(ppi)
out["ecc_oracle_s"] = time.perf_counter() - t
out["ecc_nodes"], out["ecc_edges"] = a.nodes, int(ppi.nnz)
out["ecc_oracle_equals_reference"] = bool(np.array_equal(e_ref.row, e_or.row) and np.array_equal(e_ref.col, e_or.col)
                                          and np.array_equal(e_ref.data, e_or.data))

n = a.dense_nodes
ppi = synth.ppi_problem(n, int(a.edges * n / a.nodes), "normal", 70, feat_dims=(3, 4, 4)).scipy_ppi()
rng = np.random.default_rng(1)
nor, inter = (coo_matrix(rng.random((n, n)) * 2 - 1) for _ in range(2))
t = time.perf_counter()
m_ref = ref.modify_network_topology(ppi, nor, inter, 2)
out["rewire_reference_s"] = time.perf_counter() - t
m_or = po.modify_network_topology(ppi, nor, inter, 2)
out["rewire_nodes"] = n
out["rewire_oracle_equals_reference"] = bool(np.array_equal(m_ref.row, m_or.row) and np.array_equal(m_ref.col, m_or.col))
print(json.dumps(out))
