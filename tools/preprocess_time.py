"""Times the preprocessing kernels at the PPI shape (24 041 nodes, 1.4 M directed edges) with CUDA events:
    python tools/preprocess_time.py [--nodes N --edges E]
ECC: whole call on device-resident COO ids (two sorts + intersection kernel).  Rewiring: moments + decision + emission on
two dense float64 N x N matrices generated on the device (2 x 4.6 GB at N = 24 041)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plagnn_b200 import preprocess as pp, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=24041)
ap.add_argument("--edges", type=int, default=1400000)
ap.add_argument("--once", action="store_true", help="one call of each step, no timing loop (for an ncu capture)")
a = ap.parse_args()
dev = torch.device("cuda:0")
ppi = synth.ppi_problem(a.nodes, a.edges, "normal", 70, feat_dims=(3, 4, 4)).scipy_ppi()
row = torch.from_numpy(ppi.row.astype(np.int32)).to(dev)
col = torch.from_numpy(ppi.col.astype(np.int32)).to(dev)


def timed(fn, reps=3):
    if a.once:
        fn()
        torch.cuda.synchronize()
        return None
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


out = {"nodes": a.nodes, "edges": int(ppi.nnz)}
out["ecc_ms"] = timed(lambda: pp.ecc_device(row, col, a.nodes, 0.0))
gen = torch.Generator(device=dev).manual_seed(1)
nor = torch.rand((a.nodes, a.nodes), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
inter = torch.rand((a.nodes, a.nodes), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
expr = torch.abs(torch.randn((a.nodes, 3), dtype=torch.float64, device=dev, generator=gen) * 2 + 8)
out["pearson_ms"] = timed(lambda: pp.pearson_matrix(expr, dev))
out["moments_ms"] = timed(lambda: pp.diff_moments(nor, inter))
mean, std = pp.diff_moments(nor, inter)
out["rewire_ms"] = timed(lambda: pp.rewire_device(row, col, a.nodes, nor, inter, mean - 2 * std, mean + 2 * std))
if a.once:
    print(json.dumps(out))
    sys.exit(0)
gb = 2 * 8 * a.nodes * a.nodes / 1e9
out["moments_GBps"] = 2 * gb / (out["moments_ms"] * 1e-3)           # two passes over both matrices
out["rewire_GBps"] = gb / (out["rewire_ms"] * 1e-3)                 # one pass (+ the bit matrices)
out["pearson_write_GBps"] = 8 * a.nodes * a.nodes / 1e9 / (out["pearson_ms"] * 1e-3)   # the N x N float64 matrix written once
print(json.dumps(out))
