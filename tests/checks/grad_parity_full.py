"""Full-size (N = 24 041) gradient parity: GPU vs the fp32 oracle and vs the float64 oracle, per parameter tensor."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import plagnn_b200 as P
from plagnn_b200 import synth
from oracle import plagnn_oracle as orc
from tests.helpers import copy_params, rel_err

cuda = torch.device("cuda:0")
prob = synth.ppi_problem(state="inter")
n = prob.num_nodes
ids = list(range(n))
g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(cuda)
go = orc.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids)
torch.manual_seed(70)
mo = orc.GNN32Ref(503, 400, 300, 200, 100, 12)
m = P.GNN32(503, 400, 300, 200, 100, 12)
copy_params(m, mo)
m = m.to(cuda)
md = copy.deepcopy(mo).double()
w = orc.weight_cal(prob.loc)
idx = [int(i) for i in prob.labelled[::2]]
lo = mo(go, go.ndata["feat"]); loss_o = orc.multi_loss(lo[idx], go.ndata["loc"][idx], w); loss_o.backward()
ld = md(go, go.ndata["feat"].double()); loss_d = orc.multi_loss(ld[idx], go.ndata["loc"][idx].double(), w); loss_d.backward()
lc = m(g, g.ndata["feat"]); loss_c = P.multi_loss_indexed(lc, g.ndata["loc"], torch.as_tensor(idx, device=cuda), w); loss_c.backward()
print(f"logits: gpu vs f32 oracle {rel_err(lc, lo):.2e}, gpu vs f64 {rel_err(lc, ld):.2e}, f32 oracle vs f64 {rel_err(lo, ld):.2e}")
print(f"loss: gpu {loss_c.item():.8f} f32 {loss_o.item():.8f} f64 {loss_d.item():.10f}")
for (name, pc), po, pd in zip(m.named_parameters(), mo.parameters(), md.parameters()):
    print(f"{name:22s} gpu-f32 {rel_err(pc.grad, po.grad):.2e}  gpu-f64 {rel_err(pc.grad, pd.grad):.2e}  f32-f64 {rel_err(po.grad, pd.grad):.2e}")
