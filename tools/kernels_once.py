"""Runs each hot kernel a few times on the PPI-shaped sizes (for ncu --set full captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth

dev = torch.device("cuda:0")
N = 24041
which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
if which in ("all", "gemm"):
    x = ops.aligned(torch.randn(N, 503, device=dev))
    w = ops.aligned(torch.randn(503, 503, device=dev) * 0.05)
    b = torch.randn(503, device=dev)
    for _ in range(3):
        ops.gemm(N, 503, [(x, 0, w, 0, 503)], bias=b, act=ops.ACT_RELU, backend=ops.GEMM_AUTO)
    dz = ops.aligned(torch.randn(N, 400, device=dev))
    out = torch.empty(400, 503, device=dev)
    for _ in range(3):
        ops.gemm(400, 503, [(dz, 1, x, 1, N)], out=out, backend=ops.GEMM_AUTO)
if which in ("all", "narrow"):
    h4 = ops.aligned(torch.randn(N, 100, device=dev))
    w2 = ops.aligned(torch.randn(12, 100, device=dev) * 0.1)
    b2 = torch.randn(12, device=dev)
    dz5 = ops.aligned(torch.randn(N, 12, device=dev))
    for _ in range(3):
        ops.gemm(N, 12, [(h4, 0, w2, 0, 100)], bias=b2, act=ops.ACT_SIGMOID)                 # narrow_n_kernel
        ops.gemm(N, 100, [(dz5, 0, w2, 1, 12)], gate=h4, gate_act=ops.ACT_LEAKY)             # narrow_k_kernel
        ops.gemm_wgrad_bias(dz5, h4)                                                         # narrow_wgrad_kernel + reduce
if which in ("all", "spmm"):
    prob = synth.ppi_problem(state="inter")
    g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=N).add_self_loop().to(dev)
    m = ops.aligned(torch.relu(torch.randn(N, 503, device=dev)))
    for _ in range(3):
        o, a = ops.spmm_max_fwd(g.csc(), m)
    dzz = ops.aligned(torch.randn(N, 503, device=dev))
    for _ in range(3):
        ops.spmm_max_bwd(dzz, a, o, N)
torch.cuda.synchronize()
print("ok")
