"""Generates tests/golden/*.npz by running the REFERENCE's own code (imported from
/root/reference/code, which exists only in the build container) on small seeded inputs.

    python tests/golden/make_golden.py

What is pinned by the reference itself (no DGL needed): multi_loss, weight_cal, protein_loc_correction,
performances_record (code/train.py) and scaling (code/main.py).  The end-to-end fixture runs the
reference's UNCHANGED train.train() loop with `dgl` replaced by the CPU oracle (DGL is not installable
here), so it pins the loop / loss / Adam / KFold / seeding behaviour around the oracle's SAGEConv.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/code"
OUT = os.path.dirname(os.path.abspath(__file__))

import plagnn_b200  # noqa: E402  (synthetic inputs only)
from plagnn_b200 import synth  # noqa: E402
from oracle import plagnn_oracle as orc  # noqa: E402


class _Nodes:
    def __init__(self, g):
        self.g = g

    def __getitem__(self, ids):
        g = self.g

        class _D:
            def __setitem__(self, k, v):
                g.ndata[k] = v
        return types.SimpleNamespace(data=_D())


def install_stubs():
    dgl = types.ModuleType("dgl")

    def graph(data, num_nodes=None):
        s, d = data
        g = orc.OracleGraph(np.asarray(s), np.asarray(d), num_nodes)
        g.nodes = _Nodes(g)
        return g

    def add_self_loop(g):
        s, d = orc.add_self_loop(g.src, g.dst, g.num_nodes)
        g2 = orc.OracleGraph(s, d, g.num_nodes)
        g2.nodes = _Nodes(g2)
        return g2

    dgl.graph, dgl.add_self_loop, dgl.seed = graph, add_self_loop, (lambda s: None)
    nnm, pt = types.ModuleType("dgl.nn"), types.ModuleType("dgl.nn.pytorch")
    pt.SAGEConv = orc.SAGEConvPoolRef
    nnm.pytorch = pt
    dgl.nn = nnm
    sys.modules.update({"dgl": dgl, "dgl.nn": nnm, "dgl.nn.pytorch": pt})
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


def main():
    install_stubs()
    tmp = tempfile.mkdtemp(prefix="plagnn_golden_")
    N, E, DIMS = 300, 3000, (3, 20, 20)
    # tiny tree in the reference's on-disk formats
    normal, inter = synth.write_reference_tree(tmp, "GSE74572", N, E, seed=70)
    # feature blocks are written full width by write_reference_tree; re-write the narrow ones used here
    gm = os.path.join(tmp, "data", "generate_materials")
    ds = os.path.join(gm, "GSE74572_data")
    np.save(os.path.join(gm, "ECC_normal_pca.npy"), normal.ecc[:, :DIMS[2]])
    np.save(os.path.join(ds, "GCN_normal_pca.npy"), normal.gcn[:, :DIMS[1]])
    os.chdir(os.path.join(tmp, "code"))
    import statistics  # noqa: F401  stdlib module; the reference ships a same-named script that must not shadow it
    import torch._dynamo  # noqa: F401  (imports stdlib `statistics` lazily otherwise)
    sys.path.append(REF)
    import train as ref_train      # the reference's train.py, unmodified
    import utils as ref_utils      # the reference's utils.py, unmodified (dgl -> oracle stub)
    # main.py runs its pipeline at import time; take only its `scaling` function, verbatim, via ast
    import ast
    src = open(os.path.join(REF, "main.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "scaling"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REF, "main.py"), "exec"), ns)
    ref_main = types.SimpleNamespace(scaling=ns["scaling"])

    rng = np.random.default_rng(5)
    # ---- (1) multi_loss incl. saturated probabilities and its gradient ---------------------------
    p = rng.uniform(0.0, 1.0, size=(64, 12)).astype(np.float32)
    p[0, :4] = [0.0, 1.0, 1e-10, 1.0 - 1e-7]
    p[1, :4] = [1e-9, 5e-10, 0.99999994, 2e-9]
    t = (rng.uniform(size=(64, 12)) < 0.3).astype(np.float32)
    w = rng.uniform(0.5, 40.0, size=12)            # float64, like weight_cal's output
    pt_ = torch.tensor(p, requires_grad=True)
    loss = ref_train.multi_loss(pt_, torch.tensor(t), w)
    loss.backward()
    # ---- (2) weight_cal ------------------------------------------------------------------------
    wc = ref_train.weight_cal(normal.loc)
    # ---- (3) label decision + metrics ----------------------------------------------------------
    probs = torch.tensor(rng.uniform(0.01, 0.99, size=(200, 12)).astype(np.float32))
    pred = ref_train.protein_loc_correction(probs, 0.1)
    truth = torch.tensor((rng.uniform(size=(200, 12)) < 0.2).astype(np.float32))
    truth[:, 0] = torch.where(truth.sum(1) == 0, torch.ones(200), truth[:, 0])
    aim, cov, acc = ref_train.performances_record(truth, pred)
    # ---- (4) scaling ---------------------------------------------------------------------------
    sc_in = rng.uniform(0.01, 0.99, size=(50, 12))
    sc_out = ref_main.scaling(sc_in)
    np.savez_compressed(os.path.join(OUT, "reference_functions.npz"),
                        loss_p=p, loss_t=t, loss_w=w, loss_value=loss.detach().numpy(), loss_grad=pt_.grad.numpy(),
                        wc_loc=normal.loc, wc_out=wc,
                        lc_probs=probs.numpy(), lc_alpha=0.1, lc_pred=pred.numpy(),
                        pr_truth=truth.numpy(), pr_out=np.array([aim, cov, acc]),
                        sc_in=sc_in, sc_out=sc_out)

    # ---- (5) the unchanged train loop on the tiny tree (oracle SAGEConv underneath) -------------
    import random
    from scipy.sparse import load_npz
    seed = 70                                       # main_normal.py:11-16
    random.seed(seed); torch.manual_seed(seed); np.random.seed(seed)
    ppi = load_npz(os.path.join(gm, "PPI_normal.npz"))
    ecc = np.load(os.path.join(gm, "ECC_normal_pca.npy"))
    gcn = np.load(os.path.join(ds, "GCN_normal_pca.npy"))
    loc = load_npz(os.path.join(gm, "loc_matrix.npz"))
    expr = np.load(os.path.join(ds, "expr_normal.npy"))
    with open(os.path.join(gm, "protein_ppi.json")) as f:
        uniprot = json.load(f)
    g = ref_utils.create_graph(ppi, ecc, gcn, loc, expr, uniprot)
    log_path = os.path.join(tmp, "data", "log", "GSE74572", "normal") + "/"
    os.makedirs(log_path, exist_ok=True)
    FOLDS, EPOCHS = 2, 3
    ref_train.train(g, lr=5e-5, fold_num=FOLDS, epoch_num=EPOCHS, alpha_list=[0.1], device="cpu", path=log_path)
    with open(log_path + "fig_data_1.json") as f:
        fig = json.load(f)
    tl = np.array([fig["train"]["0.1"][str(k)]["loss"] for k in range(1, FOLDS + 1)])
    vl = np.array([fig["validation"]["0.1"][str(k)]["loss"] for k in range(1, FOLDS + 1)])
    logits = np.stack([np.load(log_path + f"1_{k}_loc_logits.npy") for k in range(1, FOLDS + 1)])
    csc = g.csc()
    np.savez_compressed(os.path.join(OUT, "train_loop_tiny.npz"),
                        ppi_row=ppi.row.astype(np.int32), ppi_col=ppi.col.astype(np.int32),
                        ecc=ecc, gcn=gcn, expr=expr, loc=loc.toarray(), labelled=normal.labelled,
                        indptr=csc[0], indices=csc[1], eids=csc[2],
                        train_loss=tl, val_loss=vl, logits=logits,
                        folds=FOLDS, epochs=EPOCHS, lr=5e-5, seed=seed, fold_seed=12)
    print("written:", sorted(os.listdir(OUT)))
    print("train losses (round 1):", tl)


if __name__ == "__main__":
    main()
