"""Aggregation over the WHOLE 1 M-node / 100 M-edge graph on a column slice of F/P columns (what one GPU of the feature
partition runs): weighted sum and max reducer at F/P = 32, 64, 128, 256, event-timed.  Run once with PLAGNN_SPMM_NARROW=0
and once with =1 to compare the wide kernel with the sub-warp kernel on the narrow slices."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth

dev = torch.device("cuda:0")
n, e = 1_000_000, 100_000_000
sg = synth.scaled_graph(n, e, seed=1234, device=dev)
csc = P.build_csr(sg.dst.to(torch.int32), sg.src.to(torch.int32), n, False)
w = sg.weight[csc.eids.long()].contiguous()
scale = (1.0 / csc.degrees.clamp(min=1).float()).contiguous()
src32, dst32, wfull = sg.src.to(torch.int32), sg.dst.to(torch.int32), sg.weight
del sg


def slab_structs(n_slabs):
    rows_per = (n + n_slabs - 1) // n_slabs
    sid = src32 // rows_per
    cs, ws = [], []
    for i in range(n_slabs):
        m = sid == i
        c = P.build_csr(dst32[m], src32[m], n, False)
        cs.append(c)
        ws.append(wfull[m][c.eids.long()].contiguous())
    return cs, ws


out = {"narrow": os.environ.get("PLAGNN_SPMM_NARROW", "1"), "narrow_u": os.environ.get("PLAGNN_SPMM_NARROW_U", "4"),
       "edges": int(csc.num_edges)}
SLABS = os.environ.get("PLAGNN_TIME_SLABS", "0") == "1"
for f in ((32, 64) if os.environ.get('PLAGNN_TIME_NARROW_ONLY') == '1' else (32, 64, 128, 256)):
    x = ops.alloc(n, f, dev)
    x.copy_(torch.randn(n, f, device=dev))
    bias = torch.zeros(f, device=dev)
    for name, fn in (("sum", lambda: ops.spmm_sum(csc, x, w=w, scale=scale, bias=bias, act=ops.ACT_LEAKY, w_in_csr_order=True)),
                     ("max", lambda: ops.spmm_max_fwd(csc, x))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            fn()
        t.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(t) / 5
        alg = 4 * f * csc.num_edges + 4 * csc.num_edges * (2 if name == "sum" else 1) + 4 * f * n * (1 if name == "sum" else 2)
        out[f"{name}/{f}"] = {"ms": round(ms, 3), "algorithmic_gb_per_s": round(alg / ms / 1e6, 0)}
    if SLABS and f <= 64 and os.environ.get("PLAGNN_SPMM_NARROW", "1") != "0":
        for n_slabs in ((2, 4, 8) if f == 32 else (4, 8, 16)):
            cs, ws = slab_structs(n_slabs)
            for name, fn in (("sum", lambda: ops.spmm_sum_slabs(cs, x, ws=ws, scale=scale, bias=bias, act=ops.ACT_LEAKY)),
                             ("max", lambda: ops.spmm_max_slabs(cs, x))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(5):
                    fn()
                t.record()
                torch.cuda.synchronize()
                out[f"{name}/{f}/slabs{n_slabs}"] = {"ms": round(s.elapsed_time(t) / 5, 3), "slab_mb": round(n * f * 4 / n_slabs / 2 ** 20, 1)}
            del cs, ws
    del x
print(json.dumps(out))
