"""Shared test helpers (tests are the only product-side code allowed to import oracle/)."""
import numpy as np
import torch

from oracle import plagnn_oracle as orc

REL_TOL = 1e-5   # north_star: fp32 logits, gradients and per-epoch loss within 1e-5 relative


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error against the reference b)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    denom = b.abs().max().item()
    return ((a - b).abs().max().item() / denom) if denom > 0 else (a - b).abs().max().item()


def random_multigraph(n, e, seed, hubs=0, hub_deg=0, isolated=0):
    """COO with duplicates, optional high in-degree hubs and rows without in-edges."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, size=e)
    dst = rng.integers(isolated, n, size=e) if isolated else rng.integers(0, n, size=e)
    if hubs:
        hs = rng.integers(0, n, size=hubs * hub_deg)
        hd = np.repeat(rng.choice(np.arange(isolated, n), size=hubs, replace=False), hub_deg)
        src, dst = np.concatenate([src, hs]), np.concatenate([dst, hd])
        perm = rng.permutation(len(src))
        src, dst = src[perm], dst[perm]
    return src.astype(np.int32), dst.astype(np.int32)


def copy_params(dst_model, src_model):
    """Copies parameters by name (the drop-in and the oracle share DGL's SAGEConv parameter names)."""
    sd = {k: v.detach().clone() for k, v in src_model.state_dict().items()}
    missing = dst_model.load_state_dict(sd, strict=True)
    return missing


def oracle_graph_from(src, dst, n, self_loops=True):
    s, d = (orc.add_self_loop(src.astype(np.int64), dst.astype(np.int64), n) if self_loops
            else (src.astype(np.int64), dst.astype(np.int64)))
    return orc.OracleGraph(s, d, n)
