import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plagnn_b200 as P
from plagnn_b200 import ops, synth
from tools.microbench import timeit
dev = torch.device("cuda:0")
prob = synth.ppi_problem(state="inter")
N = prob.num_nodes
for chunk in (64, 128, 256, 512, 1024, 4096, 1 << 20):
    g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=N)
    g.chunk = chunk
    g = g.add_self_loop().to(dev)
    csc = g.csc()
    row = f"chunk={chunk:8d} items/hubs/slots={csc.counts} "
    for f in (503, 400, 300):
        x = ops.aligned(torch.relu(torch.randn(N, f, device=dev)))
        ms = timeit(lambda: ops.spmm_max_fwd(csc, x), reps=30)
        row += f"| F={f}: {ms:.4f} ms "
    print(row, flush=True)
