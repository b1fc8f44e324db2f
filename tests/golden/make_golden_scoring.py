"""Golden vectors for the alteration-scoring step, produced by the REFERENCE's own code (run in the build container,
where /root/reference exists):  `scaling` and the first statements of `misloc_protein_record` are taken verbatim from
code/main.py through ast (the module itself runs its whole pipeline at import time and reads data files that are not
in the repository); `mat_merge`'s accumulation loop (code/main.py:41-47) is restated on in-memory matrices.

    python tests/golden/make_golden_scoring.py        -> tests/golden/scoring.npz
"""
import ast
import os

import numpy as np

REF = "/root/reference/code"
OUT = os.path.dirname(os.path.abspath(__file__))

src = open(os.path.join(REF, "main.py")).read()
tree = ast.parse(src)
fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
ns = {"np": np}
exec(compile(ast.Module(body=[fns["scaling"]], type_ignores=[]), "main.py", "exec"), ns)
scaling = ns["scaling"]

# the scoring statements of misloc_protein_record: everything after loc_map up to (and including) diff_indices.reverse()
body = fns["misloc_protein_record"].body
stmts = []
for st in body:
    seg = ast.get_source_segment(src, st) or ""
    if seg.startswith(("normal = scaling", "inter = scaling", "diff_matrix =", "diff_indices =", "diff_indices.reverse")):
        stmts.append(st)
assert len(stmts) == 5, [ast.get_source_segment(src, s) for s in stmts]


def ref_scores(normal_mat, inter_mat):
    env = {"np": np, "scaling": scaling, "normal_mat": normal_mat, "inter_mat": inter_mat}
    exec(compile(ast.Module(body=stmts, type_ignores=[]), "main.py", "exec"), env)
    return env["normal"], env["inter"], env["diff_matrix"], np.array(env["diff_indices"], dtype=np.int64)


rng = np.random.default_rng(2024)
# (1) scaling on float32 (what np.load returns for a run's logits) and float64
sc32_in = rng.uniform(0.001, 0.999, size=(97, 12)).astype(np.float32)
sc32_out = scaling(sc32_in)
sc64_in = rng.uniform(0.001, 0.999, size=(97, 12))
sc64_out = scaling(sc64_in)
# (2) mat_merge: 100 runs (code/main.py:41-47: mat_cnt = zeros; mat_cnt += scaling(mat); mat_cnt /= 100)
runs = rng.uniform(0.001, 0.999, size=(100, 64, 12)).astype(np.float32)
mat_cnt = np.zeros((64, 12))
for mat in runs:
    mat = scaling(mat)
    mat_cnt += mat
mat_cnt /= 100
# (3) scores and ranking (the scaled matrices contain exact zeros, so diff holds +-inf / nan, as in the real pipeline)
normal_mat = rng.uniform(0.01, 0.99, size=(80, 12))
inter_mat = normal_mat * rng.uniform(0.8, 1.25, size=(80, 12))
inter_mat[3] = normal_mat[3]                      # a protein whose scores do not change (ties at 0 after scaling? not exactly)
with np.errstate(divide="ignore", invalid="ignore"):
    n_s, i_s, diff, order = ref_scores(normal_mat, inter_mat)
np.savez_compressed(os.path.join(OUT, "scoring.npz"), sc32_in=sc32_in, sc32_out=sc32_out, sc64_in=sc64_in, sc64_out=sc64_out,
                    runs=runs, merged=mat_cnt, normal_mat=normal_mat, inter_mat=inter_mat, normal=n_s, inter=i_s, diff=diff,
                    order=order)
print("wrote scoring.npz; nan", int(np.isnan(diff).sum()), "inf", int(np.isinf(diff).sum()), "dtype", sc32_out.dtype)
