"""Is the epoch host-bound?  Host enqueue time per epoch (no sync inside the loop) vs device time (CUDA events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import plagnn_b200 as P
from plagnn_b200 import synth
dev = torch.device("cuda:0")
prob = synth.ppi_problem(state="inter")
n = prob.num_nodes
g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, list(range(n))).to(dev)
feat, lab = g.ndata["feat"], g.ndata["loc"]
model = P.GNN32(503, 400, 300, 200, 100, 12).to(dev)
opt = P.FusedAdam(model.parameters(), lr=5e-5)
w = P.weight_cal(prob.loc)
idx = torch.as_tensor(prob.labelled[::2], device=dev)
def epoch():
    opt.zero_grad()
    logits = model(g, feat)
    loss = P.multi_loss_indexed(logits, lab, idx, w)
    loss.backward()
    opt.step()
for _ in range(10): epoch()
torch.cuda.synchronize()
for reps in (50, 200):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(reps): epoch()
    t_enq = time.perf_counter() - t0
    e.record(); torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{reps} epochs: host enqueue {t_enq / reps * 1e3:.3f} ms/epoch, wall {t_all / reps * 1e3:.3f} ms/epoch, device (events) {s.elapsed_time(e) / reps:.3f} ms/epoch")
# split of the host time
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): epoch()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
