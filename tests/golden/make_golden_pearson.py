"""Golden vectors for the Pearson co-expression matrix, produced by the REFERENCE's own statements.

`construct_gcn_matrix` (code/data_preprocess.py:128-172) cannot run as a whole in this image: its table-filling loop assigns
float rows into an int64 DataFrame, which pandas 3 refuses (the reference pins an older pandas).  Its numeric tail — from
`expr_pcc = np.corrcoef(expr_gcn)` to `gcn = coo_matrix(expr_pcc)` (:165-170) — is taken verbatim from the file through ast
and executed on an expression matrix shaped like the one the loop builds (duplicated probes averaged, PPI proteins without
expression = all-zero rows, a constant row).  Run in the build container:

    python tests/golden/make_golden_pearson.py        -> tests/golden/pearson.npz
"""
import ast
import os

import numpy as np
from scipy.sparse import coo_matrix

REF = "/root/reference/code"
OUT = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(REF, "data_preprocess.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "construct_gcn_matrix")
stmts = [st for st in fn.body if (ast.get_source_segment(src, st) or "").startswith(
    ("expr_pcc = np.corrcoef", "np.fill_diagonal(expr_pcc", "pcc_nan =", "expr_pcc[pcc_nan]", "gcn = coo_matrix"))]
assert len(stmts) == 5, [ast.get_source_segment(src, s) for s in stmts]

rng = np.random.default_rng(991)
n = 120
expr_gcn = np.zeros((n, 3))
for i in range(n):
    if rng.random() < 0.12:
        continue                                                     # no expression: the all-zero row the reference leaves
    probes = rng.normal(8.0, 2.0, size=(int(rng.integers(1, 4)), 3))
    expr_gcn[i] = probes.mean(axis=0)                                # groupby(...).agg('mean') of :151
expr_gcn[5] = 7.25                                                   # constant, non-zero: zero variance as well
env = {"np": np, "coo_matrix": coo_matrix, "expr_gcn": expr_gcn.copy()}
with np.errstate(invalid="ignore", divide="ignore"):
    exec(compile(ast.Module(body=stmts, type_ignores=[]), "data_preprocess.py", "exec"), env)
gcn = env["gcn"]
np.savez_compressed(os.path.join(OUT, "pearson.npz"), expr=expr_gcn, pcc=gcn.toarray())
print("wrote pearson.npz:", expr_gcn.shape, "nnz", gcn.nnz, "zero-variance rows", int((expr_gcn.std(1) == 0).sum()))
