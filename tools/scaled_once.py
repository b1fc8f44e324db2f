"""One launch of each aggregation kernel on the 1 M-node / 100 M-edge graph (for ncu --set full: DRAM bytes of the kernels the
>= 70 % of HBM target is judged on): weighted sum and max at 256 columns (one GPU / the row partition), weighted sum and max at
32 columns (one rank's slice of the feature partition at 8 GPUs), the max reducer's scatter backward at 256 columns."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for f in (256, 32):
    x = ops.alloc(n, f, dev)
    x.copy_(torch.randn(n, f, device=dev))
    bias = torch.zeros(f, device=dev)
    for _ in range(reps):
        ops.spmm_sum(csc, x, w=w, scale=scale, bias=bias, act=ops.ACT_LEAKY, w_in_csr_order=True)
        o, a = ops.spmm_max_fwd(csc, x)
    if f == 256:
        dz = ops.alloc(n, f, dev)
        dz.copy_(torch.randn(n, f, device=dev))
        for _ in range(reps):
            ops.spmm_max_bwd(dz, a, o, n)
        del dz
    del x, o, a
torch.cuda.synchronize()
print("ok")
