"""Times the max aggregation (forward and scatter backward) on the PPI-shaped graph at the three layer widths."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth

dev = torch.device("cuda:0")
N = 24041
prob = synth.ppi_problem(state="inter")
g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=N).add_self_loop().to(dev)
csc = g.csc()
torch.manual_seed(0)
out = {}
for F in (503, 400, 300):
    m = ops.aligned(torch.relu(torch.randn(N, F, device=dev)))
    dz = ops.aligned(torch.randn(N, F, device=dev))
    for _ in range(5):
        o, a = ops.spmm_max_fwd(csc, m)
        ops.spmm_max_bwd(dz, a, o, N)
    reps = 40
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        o, a = ops.spmm_max_fwd(csc, m)
    e[1].record()
    for _ in range(reps):
        ops.spmm_max_bwd(dz, a, o, N)
    e[2].record()
    torch.cuda.synchronize()
    out[F] = (round(e[0].elapsed_time(e[1]) / reps, 4), round(e[1].elapsed_time(e[2]) / reps, 4))
print(os.environ.get("PLAGNN_LIB_PATH", "default"), out)
