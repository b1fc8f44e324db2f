"""
Synthetic stand-ins for the PLA-GNN inputs (the real ``data/PPI.7z`` is absent from the
reference tree, ``.MISSING_LARGE_BLOBS:1``).  Shapes and on-disk formats follow what the
reference reads (``code/main_normal.py:57-63``, ``code/main_inter.py:57-61``,
``code/train.py:128-129,151-155``) and what ``code/data_preprocess.py`` writes
(``:273-277,303-306,324,447-454,531-546``); see SURVEY.md §8(d).

Everything here is plain torch/numpy (device-agnostic): it feeds the CUDA path, the oracle and
the benchmark with identical inputs.  No kernels live in this file.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np
import torch

PPI_NODES = 24041        # code/main.py:40, code/performance.py:99
PPI_CLASSES = 12         # data/support_materials/cellular_component.txt:1-12
PPI_FEATS = 503          # 3 expr + 250 co-expression PCA + 250 ECC PCA (code/utils.py:47-48)
PPI_EDGES = 1_400_000    # directed; estimate, the real count is not in the repository
GO_TERMS = ["GO:0005938", "GO:0005829", "GO:0015629", "GO:0005794", "GO:0005783", "GO:0005730",
            "GO:0005777", "GO:0005739", "GO:0005764", "GO:0005813", "GO:0005886", "GO:0005654"]


def powerlaw_edges(num_nodes: int, num_directed: int, exponent: float = 2.2, seed: int = 70,
                   max_degree: int | None = None, device="cpu"):
    """Chung-Lu style undirected simple graph, returned symmetrised as directed (src, dst)
    int64 tensors with exactly ``num_directed`` entries (rounded down to even), zero diagonal,
    no duplicate pairs, every node with degree >= 1."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    half = num_directed // 2
    ranks = torch.arange(1, num_nodes + 1, device=device, dtype=torch.float64)
    w = ranks.pow(-1.0 / (exponent - 1.0))
    if max_degree is not None:
        # cap expected degree: expected deg_i = 2*half*w_i/sum(w)
        for _ in range(8):
            cap = max_degree * w.sum() / (2.0 * half)
            w = torch.minimum(w, cap)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    # every node gets one partner first (min degree 1): node i <-> random weighted partner
    base_a = torch.arange(num_nodes, device=device)
    base_b = torch.searchsorted(cdf, torch.rand(num_nodes, generator=gen, device=device, dtype=torch.float64))
    base_b = base_b.clamp_(max=num_nodes - 1)
    clash = base_b == base_a
    base_b[clash] = (base_a[clash] + 1) % num_nodes
    keys = _pair_keys(base_a, base_b, num_nodes)
    keys = torch.unique(keys)
    while keys.numel() < half:
        need = int((half - keys.numel()) * 1.3) + 1024
        a = torch.searchsorted(cdf, torch.rand(need, generator=gen, device=device, dtype=torch.float64))
        b = torch.searchsorted(cdf, torch.rand(need, generator=gen, device=device, dtype=torch.float64))
        a.clamp_(max=num_nodes - 1)
        b.clamp_(max=num_nodes - 1)
        ok = a != b
        keys = torch.unique(torch.cat([keys, _pair_keys(a[ok], b[ok], num_nodes)]))
    if keys.numel() > half:
        # drop random surplus pairs, but never the ones that guarantee min degree
        base = torch.unique(_pair_keys(base_a, base_b, num_nodes))
        is_base = torch.isin(keys, base)
        extra = keys[~is_base]
        keep_extra = half - int(is_base.sum())
        perm = torch.randperm(extra.numel(), generator=gen, device=device)[:max(keep_extra, 0)]
        keys = torch.cat([keys[is_base], extra[perm]])
    lo = keys // num_nodes
    hi = keys % num_nodes
    src = torch.cat([lo, hi])
    dst = torch.cat([hi, lo])
    return src, dst


def _pair_keys(a, b, n):
    lo = torch.minimum(a, b).to(torch.int64)
    hi = torch.maximum(a, b).to(torch.int64)
    return lo * n + hi


def coo_order(src, dst, form: str, seed: int, num_nodes: int):
    """'normal': entries shuffled (PPI_normal.npz is built from a Python set,
    data_preprocess.py:83-108); 'inter': row-major order (coo_matrix(dense), :253)."""
    if form == "normal":
        gen = torch.Generator(device=src.device)
        gen.manual_seed(seed + 1)
        perm = torch.randperm(src.numel(), generator=gen, device=src.device)
    elif form == "inter":
        perm = torch.argsort(src * num_nodes + dst)
    else:
        raise ValueError(form)
    return src[perm], dst[perm]


def rewire(src, dst, num_nodes: int, frac: float = 0.005, seed: int = 7):
    """Perturbation-state adjacency: drop ``frac`` of the undirected pairs and add as many new
    ones (stand-in for data_preprocess.modify_network_topology, :217-257)."""
    gen = torch.Generator(device=src.device)
    gen.manual_seed(seed)
    keys = torch.unique(_pair_keys(src, dst, num_nodes))
    k = int(keys.numel() * frac)
    keep = torch.randperm(keys.numel(), generator=gen, device=src.device)[k:]
    keys = keys[keep]
    target = keys.numel() + k
    while keys.numel() < target:
        a = torch.randint(0, num_nodes, (2 * k + 16,), generator=gen, device=src.device)
        b = torch.randint(0, num_nodes, (2 * k + 16,), generator=gen, device=src.device)
        ok = a != b
        keys = torch.unique(torch.cat([keys, _pair_keys(a[ok], b[ok], num_nodes)]))
    keys = keys[:target]
    lo, hi = keys // num_nodes, keys % num_nodes
    return torch.cat([lo, hi]), torch.cat([hi, lo])


def node_features(num_nodes: int, seed: int = 70):
    """expr f64 N x 3 ~ |N(8,2)|, gcn / ecc f64 N x 250 with PCA-like decaying column scales."""
    rng = np.random.default_rng(seed)
    expr = np.abs(rng.normal(8.0, 2.0, size=(num_nodes, 3)))
    sig = 4.0 / np.sqrt(1.0 + np.arange(250))
    gcn = rng.normal(0.0, 1.0, size=(num_nodes, 250)) * sig
    ecc = rng.normal(0.0, 1.0, size=(num_nodes, 250)) * sig
    return expr, gcn, ecc


def labels(num_nodes: int, num_classes: int = PPI_CLASSES, labelled_frac: float = 0.4, seed: int = 70):
    """~40 % of the nodes carry 1-3 of the 12 classes; every class non-empty (weight_cal divides
    by the class count, train.py:124).  Returns (loc dense f64 N x C, labelled row indices)."""
    rng = np.random.default_rng(seed + 3)
    loc = np.zeros((num_nodes, num_classes), dtype=np.float64)
    n_lab = max(num_classes, int(num_nodes * labelled_frac))
    rows = np.sort(rng.choice(num_nodes, size=n_lab, replace=False))
    prior = rng.dirichlet(np.full(num_classes, 2.0))
    for i, r in enumerate(rows):
        if i < num_classes:
            loc[r, i] = 1.0   # guarantees every class appears
        k = rng.integers(1, 4)
        cls = rng.choice(num_classes, size=k, replace=False, p=prior)
        loc[r, cls] = 1.0
    return loc, rows.astype(np.int64)


@dataclass
class PPIProblem:
    """One (dataset, state) training input, host side, in the reference's own types."""
    ppi_row: np.ndarray      # int32 [E]   (scipy COO .row)
    ppi_col: np.ndarray      # int32 [E]
    expr: np.ndarray         # f64 [N,3]
    gcn: np.ndarray          # f64 [N,250]
    ecc: np.ndarray          # f64 [N,250]
    loc: np.ndarray          # f64 [N,12]
    labelled: np.ndarray     # int64 indices with >=1 label
    num_nodes: int

    @property
    def features(self) -> np.ndarray:
        # utils.py:47-48: hstack((expr, hstack((gcn, ecc)))) cast to float32
        return np.hstack((self.expr, np.hstack((self.gcn, self.ecc)))).astype(np.float32)

    def scipy_ppi(self):
        from scipy.sparse import coo_matrix
        n = self.num_nodes
        return coo_matrix((np.ones(len(self.ppi_row), dtype=np.int64), (self.ppi_row, self.ppi_col)), shape=(n, n))

    def scipy_loc(self):
        from scipy.sparse import coo_matrix
        return coo_matrix(self.loc)


def ppi_problem(num_nodes: int = PPI_NODES, num_directed: int = PPI_EDGES, state: str = "normal",
                seed: int = 70, feat_dims=(3, 250, 250)) -> PPIProblem:
    """PPI-shaped problem.  ``state`` 'normal' = control adjacency in
    shuffled COO order; 'inter' = rewired adjacency in row-major COO order."""
    src, dst = powerlaw_edges(num_nodes, num_directed, 2.2, seed)
    if state == "inter":
        src, dst = rewire(src, dst, num_nodes, 0.005, seed + 11)
    src, dst = coo_order(src, dst, state, seed, num_nodes)
    expr, gcn, ecc = node_features(num_nodes, seed if state == "normal" else seed + 5)
    expr, gcn, ecc = expr[:, :feat_dims[0]], gcn[:, :feat_dims[1]], ecc[:, :feat_dims[2]]
    loc, labelled = labels(num_nodes, PPI_CLASSES, 0.4, seed)
    return PPIProblem(src.numpy().astype(np.int32), dst.numpy().astype(np.int32), expr, gcn, ecc, loc,
                      labelled, num_nodes)


def write_reference_tree(root: str, gse: str = "GSE74572", num_nodes: int = 600, num_directed: int = 6000,
                         seed: int = 70):
    """Write ``<root>/data/generate_materials/...`` exactly as main_normal.py / main_inter.py /
    train.py expect to read it, so the UNCHANGED reference scripts run on synthetic inputs."""
    from scipy.sparse import save_npz
    gm = os.path.join(root, "data", "generate_materials")
    ds = os.path.join(gm, f"{gse}_data")
    os.makedirs(ds, exist_ok=True)
    os.makedirs(os.path.join(root, "data", "log"), exist_ok=True)
    os.makedirs(os.path.join(root, "code"), exist_ok=True)
    normal = ppi_problem(num_nodes, num_directed, "normal", seed)
    inter = ppi_problem(num_nodes, num_directed, "inter", seed)
    save_npz(os.path.join(gm, "PPI_normal.npz"), normal.scipy_ppi())
    np.save(os.path.join(gm, "ECC_normal_pca.npy"), normal.ecc)
    np.save(os.path.join(ds, "GCN_normal_pca.npy"), normal.gcn)
    np.save(os.path.join(ds, "expr_normal.npy"), normal.expr)
    save_npz(os.path.join(ds, "PPI_inter.npz"), inter.scipy_ppi())
    np.save(os.path.join(ds, "ECC_inter_pca.npy"), inter.ecc)
    np.save(os.path.join(ds, "GCN_inter_pca.npy"), inter.gcn)
    np.save(os.path.join(ds, "expr_inter.npy"), inter.expr)
    save_npz(os.path.join(gm, "loc_matrix.npz"), normal.scipy_loc())
    ids = sorted(f"P{i:05d}" for i in range(num_nodes))
    with open(os.path.join(gm, "protein_ppi.json"), "w") as f:
        json.dump(ids, f)
    label_list = [[ids[i], [GO_TERMS[c] for c in np.nonzero(normal.loc[i])[0]]] for i in range(num_nodes)]
    with open(os.path.join(gm, "label_list.json"), "w") as f:
        json.dump(label_list, f)
    with open(os.path.join(gm, "label_with_loc_list.json"), "w") as f:
        json.dump([int(i) for i in normal.labelled], f)
    return normal, inter


@dataclass
class ScaledGraph:
    """BASELINE.json configs[3]: power-law graph with fp32 edge weights, generated on ``device``."""
    src: torch.Tensor        # int64 [E]
    dst: torch.Tensor        # int64 [E]
    weight: torch.Tensor     # f32 [E] ~ U(0,1]
    num_nodes: int


def scaled_graph(num_nodes: int = 1_000_000, num_directed: int = 100_000_000, seed: int = 1234,
                 max_degree: int = 100_000, device="cpu") -> ScaledGraph:
    src, dst = powerlaw_edges(num_nodes, num_directed, 2.2, seed, max_degree=max_degree, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 2)
    # the generator numbers nodes by decreasing expected degree; relabel at random so that contiguous row blocks
    # (the 1-D partition) hold comparable numbers of rows AND in-edges
    relabel = torch.randperm(num_nodes, generator=gen, device=device)
    src, dst = relabel[src], relabel[dst]
    # symmetric weights: w(u,v) == w(v,u) (first half and second half of the arrays mirror each other)
    half = src.numel() // 2
    wh = 1.0 - torch.rand(half, generator=gen, device=device, dtype=torch.float32)
    return ScaledGraph(src, dst, torch.cat([wh, wh]), num_nodes)
