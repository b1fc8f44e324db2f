"""Drop-in for ``code/utils.py::create_graph`` (:28-51): same signature, same node data, graph built for
the device kernels instead of through DGL."""
from __future__ import annotations

import numpy as np
import torch as th

from .graph import add_self_loop as _add_self_loop, graph as _make_graph


def create_graph(ppi, ecc, gcn, loc, expr, uniprot):
    """ppi: scipy COO adjacency (edges row -> col); ecc / gcn / expr: dense float64 feature blocks;
    loc: scipy sparse N x C labels; uniprot: list of the N protein ids.
    Node data: 'loc' float32 N x C, 'feat' float32 N x (3+250+250) = [expr | gcn | ecc]."""
    num_nodes = len(uniprot)
    # int32 arrays go to the device as they are; no per-edge Python objects (the reference builds lists)
    g = _make_graph((np.asarray(ppi.row), np.asarray(ppi.col)), num_nodes=num_nodes)
    g = _add_self_loop(g)
    g.ndata["loc"] = th.from_numpy(loc.toarray().astype(np.float32))
    g.ndata["feat"] = th.tensor(np.hstack((expr, np.hstack((gcn, ecc)))), dtype=th.float)
    return g
