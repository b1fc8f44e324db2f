"""BASELINE.json configs[4]: feature-width and depth sweep (hidden 64-1024, 2-8 SAGEConv-pool layers) on the
PPI-shaped graph; per-epoch time split into aggregation vs dense contraction (in-library CUDA-event profile).

    python tools/sweep.py > gpurun_out/sweep.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth


class DeepSage(torch.nn.Module):
    def __init__(self, in_feats, hidden, layers, classes=12):
        super().__init__()
        dims = [in_feats] + [hidden] * layers
        self.convs = torch.nn.ModuleList([P.SAGEConv(dims[i], dims[i + 1], "pool") for i in range(layers)])
        self.out = torch.nn.Linear(hidden, classes)

    def forward(self, g, x):
        h = x
        for c in self.convs:
            h = torch.nn.functional.leaky_relu(c(g, h))
        from plagnn_b200.nn import LinearActFunction
        return LinearActFunction.apply(h, self.out.weight, self.out.bias, ops.ACT_SIGMOID)


def main():
    dev = torch.device("cuda:0")
    prob = synth.ppi_problem(state="inter")
    n = prob.num_nodes
    g = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, list(range(n))).to(dev)
    w = P.weight_cal(prob.loc)
    idx = torch.as_tensor(prob.labelled, device=dev)
    x, y = g.ndata["feat"], g.ndata["loc"]
    rows = []
    for hidden in (64, 128, 256, 512, 1024):
        for layers in (2, 4, 8):
            torch.manual_seed(0)
            model = DeepSage(x.shape[1], hidden, layers).to(dev)
            opt = P.FusedAdam(model.parameters(), lr=5e-5)

            def epoch():
                opt.zero_grad()
                loss = P.multi_loss_indexed(model(g, x), y, idx, w)
                loss.backward()
                opt.step()

            for _ in range(3):
                epoch()
            torch.cuda.synchronize()
            steps = 5
            ops.profile_start()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(steps):
                epoch()
            e.record()
            torch.cuda.synchronize()
            prof = ops.profile_stop()
            ms = s.elapsed_time(e) / steps
            gemm = sum(t for k, (c, t) in prof.items() if k[0] == "gemm") / steps
            spmm = sum(t for k, (c, t) in prof.items() if k[0].startswith("spmm")) / steps
            rows.append({"hidden": hidden, "layers": layers, "ms_per_epoch": round(ms, 3), "gemm_ms": round(gemm, 3),
                         "spmm_ms": round(spmm, 3), "spmm_share": round(spmm / (gemm + spmm), 3)})
            print(rows[-1], file=sys.stderr, flush=True)
            del model, opt
            torch.cuda.empty_cache()
    print(json.dumps({"graph": f"PPI-shaped N={n}, E'={g.csc().num_edges}, F_in={x.shape[1]}", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
