"""Aggregation kernels on the 1 M-node / 100 M-edge graph, every loop variant in ONE process (the graph is built once):
narrow kernel (32 / 64 columns: one GPU's slice of the feature partition) with PLAGNN_SPMM_NARROW_PIPE = 0..3 and
PLAGNN_SPMM_L2HINT = 0..3, wide sum kernel (128 / 256 columns) with PLAGNN_SPMM_L2HINT = 0, 1, 4, 5.  Both knobs are read
per launch by the library.  Each variant is also checked against variant 0 (the results must be bit-identical).
    python tools/spmm_variants_time.py > gpurun_out/spmm_variants.json"""
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
del sg


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / reps


out = {"edges": int(csc.num_edges)}
narrow_variants = [(8, 0), (4, 0), (2, 0), (1, 0)]
wide_variants = []
for f in (32, 64):
    x = ops.alloc(n, f, dev)
    x.copy_(torch.randn(n, f, device=dev))
    bias = torch.zeros(f, device=dev)
    ref = {}
    variants = narrow_variants if f <= 64 else [(1, h) for h in wide_variants]
    for pipe, hint in variants:
        os.environ["PLAGNN_SPMM_NARROW_WARPS"] = str(pipe)
        os.environ["PLAGNN_SPMM_L2HINT"] = str(hint)
        for name, fn in (("sum", lambda: ops.spmm_sum(csc, x, w=w, scale=scale, bias=bias, act=ops.ACT_LEAKY, w_in_csr_order=True)),
                         ("max", lambda: ops.spmm_max_fwd(csc, x))):
            if f > 64 and name == "max":
                continue            # the wide max kernel does not read the knobs
            ms = timed(fn)
            r = fn()
            r = r if isinstance(r, torch.Tensor) else r[0]
            key = f"{name}/{f}"
            same = None
            if key not in ref:
                ref[key] = r.clone()
            else:
                same = bool(torch.equal(ref[key], r))
            alg = 4 * f * csc.num_edges + 4 * csc.num_edges * (2 if name == "sum" else 1) + 4 * f * n * (1 if name == "sum" else 2)
            out[f"{key}/warps{pipe}"] = {"ms": round(ms, 3), "algorithmic_gb_per_s": round(alg / ms / 1e6, 0),
                                                    "bit_identical_to_first": same}
            print(f"{key} warps={pipe} hint={hint}: {ms:.3f} ms same={same}", file=sys.stderr, flush=True)
    del x, ref
os.environ.pop("PLAGNN_SPMM_NARROW_WARPS", None)
os.environ.pop("PLAGNN_SPMM_L2HINT", None)
print(json.dumps(out))
