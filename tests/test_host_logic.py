"""Host-side logic that needs no GPU: graph object surface, synthetic generators, class weights."""
import json
import os

import numpy as np
import torch

import plagnn_b200 as P
from plagnn_b200 import synth
from oracle import plagnn_oracle as orc


def test_graph_surface_like_the_reference_uses_it():
    g = P.graph(([0, 1, 2], [1, 2, 0]), num_nodes=4)
    g = P.add_self_loop(g)
    assert g.num_nodes() == 4 and g.num_edges() == 7
    s, d = g.edges()
    assert s.tolist() == [0, 1, 2, 0, 1, 2, 3] and d.tolist() == [1, 2, 0, 0, 1, 2, 3]
    g.nodes[list(range(4))].data["feat"] = torch.arange(8.0).reshape(4, 2)
    assert g.ndata["feat"].shape == (4, 2)
    g.nodes[[1, 3]].data["x"] = torch.ones(2, 3)
    assert g.ndata["x"].sum() == 6 and g.ndata["x"][0].sum() == 0


def test_create_graph_matches_oracle_node_data():
    prob = synth.ppi_problem(200, 1600, "normal", 3, feat_dims=(3, 10, 10))
    ids = list(range(200))
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
    assert torch.equal(g.ndata["feat"], go.ndata["feat"]) and torch.equal(g.ndata["loc"], go.ndata["loc"])
    s, d = g.edges()
    assert np.array_equal(s.numpy(), go.src) and np.array_equal(d.numpy(), go.dst)
    assert g.ndata["feat"].dtype == torch.float32 and g.ndata["feat"].shape[1] == 23


def test_powerlaw_generator_properties():
    n, e = 3000, 60000
    src, dst = synth.powerlaw_edges(n, e, 2.2, 5)
    assert src.numel() == e and (src != dst).all()
    key = src * n + dst
    assert torch.unique(key).numel() == e                     # simple graph
    assert torch.equal(torch.sort(key).values, torch.sort(dst * n + src).values)   # symmetric
    deg = torch.bincount(dst, minlength=n)
    assert deg.min() >= 1 and deg.max() > 20 * deg.float().median()   # heavy tail, no isolated node
    s2, d2 = synth.powerlaw_edges(n, e, 2.2, 5)
    assert torch.equal(src, s2) and torch.equal(dst, d2)      # seeded


def test_reference_tree_formats(tmp_path):
    normal, inter = synth.write_reference_tree(str(tmp_path), "GSE74572", 120, 900, seed=1)
    gm = tmp_path / "data" / "generate_materials"
    from scipy.sparse import load_npz
    ppi = load_npz(gm / "PPI_normal.npz")
    assert ppi.format == "coo" and ppi.dtype == np.int64 and ppi.shape == (120, 120) and (ppi.data == 1).all()
    assert np.load(gm / "ECC_normal_pca.npy").shape == (120, 250)
    assert np.load(gm / "GSE74572_data" / "expr_inter.npy").shape == (120, 3)
    lab = json.load(open(gm / "label_with_loc_list.json"))
    loc = load_npz(gm / "loc_matrix.npz").toarray()
    assert (loc[lab].sum(1) > 0).all() and (loc.sum(0) > 0).all()
    assert len(json.load(open(gm / "label_list.json"))) == 120
    inter_ppi = load_npz(gm / "GSE74572_data" / "PPI_inter.npz")
    k = inter_ppi.row.astype(np.int64) * 120 + inter_ppi.col
    assert (np.diff(k) > 0).all()                             # row-major COO order for the 'inter' form


def test_weight_cal_matches_oracle():
    loc, _ = synth.labels(500, 12, 0.4, 2)
    assert np.array_equal(P.weight_cal(loc), orc.weight_cal(loc))


def test_import_shim_and_package_dir():
    assert os.path.basename(os.path.dirname(P.__file__)) == "pla-gnn_b200"


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the package (nor the drop-ins, nor tools the product imports) may
    import it — a product path routed through the CPU restatement would void every parity claim."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "pla-gnn_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            for node in ast.walk(ast.parse(open(path).read(), path)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                if any(n == "oracle" or n.startswith("oracle.") for n in names):
                    offenders.append(os.path.relpath(path, root))
    assert not offenders, offenders


def test_fold_splits_are_the_reference_kfold():
    from sklearn.model_selection import KFold
    label = [5, 9, 11, 20, 21, 30, 41, 57, 63, 70, 88]
    got = list(__import__("plagnn_b200").pipeline.fold_splits(label, 3, 42))
    ref = list(KFold(n_splits=3, random_state=42, shuffle=True).split(label))
    assert len(got) == 3
    for (tr, va), (rt, rv) in zip(got, ref):
        assert tr.tolist() == [label[i] for i in rt] and va.tolist() == [label[i] for i in rv]    # train.py:183-188


def test_matplotlib_stand_in_yields_to_a_real_installation(tmp_path):
    """drop_in/matplotlib must not shadow a real matplotlib that sits later on sys.path."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    drop_in = os.path.join(root, "pla-gnn_b200", "drop_in")
    real = tmp_path / "site" / "matplotlib"
    real.mkdir(parents=True)
    (real / "__init__.py").write_text("REAL = True\n")
    (real / "pyplot.py").write_text("def plot():\n    return 'real plot'\n")
    code = ("import sys; sys.path[:0] = [%r, %r]\nimport matplotlib.pyplot as plt\nimport matplotlib\n"
            "print(getattr(matplotlib, 'REAL', False), plt.plot())" % (drop_in, str(tmp_path / "site")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.stdout.split() == ["True", "real", "plot"], out.stdout + out.stderr
    code = "import sys; sys.path.insert(0, %r)\nimport matplotlib.pyplot as plt\nprint('stub ok')" % drop_in
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "stub ok" in out.stdout, out.stdout + out.stderr
