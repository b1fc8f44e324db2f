"""Narrow aggregation (32 columns) on the 1 M-node / 100 M-edge graph for several plan chunk sizes (in-edges per work item)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import plagnn_b200 as P
from plagnn_b200 import ops, synth
dev = torch.device("cuda:0")
n, e = 1_000_000, 100_000_000
sg = synth.scaled_graph(n, e, seed=1234, device=dev)
dst32, src32 = sg.dst.to(torch.int32), sg.src.to(torch.int32)
out = {}
for chunk in (128, 256, 512, 1024, 2048):
    csc = P.build_csr(dst32, src32, n, False, chunk=chunk)
    w = sg.weight[csc.eids.long()].contiguous()
    scale = (1.0 / csc.degrees.clamp(min=1).float()).contiguous()
    for f in (32, 64):
        x = ops.alloc(n, f, dev); x.copy_(torch.randn(n, f, device=dev))
        bias = torch.zeros(f, device=dev)
        for name, fn in (("sum", lambda: ops.spmm_sum(csc, x, w=w, scale=scale, bias=bias, act=ops.ACT_LEAKY, w_in_csr_order=True)),
                         ("max", lambda: ops.spmm_max_fwd(csc, x))):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5): fn()
            t.record(); torch.cuda.synchronize()
            out[f"{name}/{f}/chunk{chunk}"] = round(s.elapsed_time(t) / 5, 3)
            print(name, f, chunk, out[f"{name}/{f}/chunk{chunk}"], "items", int(csc.counts[0]), "hubs", int(csc.counts[1]), file=sys.stderr, flush=True)
    del csc, w
print(json.dumps(out))
