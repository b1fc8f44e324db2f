"""Several independent models of the same fold/seed sweep trained concurrently on ONE GPU, one CUDA stream per model
(the reference trains its 100 models per condition one after the other: code/train.py:162-180).  Aggregate epochs/s.
    python tools/multi_model.py [max_models]"""
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
w = P.weight_cal(prob.loc)
rng = np.random.default_rng(0)
maxr = int(sys.argv[1]) if len(sys.argv) > 1 else 4


def make(seed):
    torch.manual_seed(seed)
    model = P.GNN32(503, 400, 300, 200, 100, 12).to(dev)
    opt = P.FusedAdam(model.parameters(), lr=5e-5)
    lab_rows = prob.labelled.copy(); rng.shuffle(lab_rows)
    idx = torch.as_tensor(np.sort(lab_rows[: len(lab_rows) * 9 // 10]), device=dev)
    return model, opt, idx, torch.cuda.Stream(device=dev)


def epoch(m):
    model, opt, idx, st = m
    with torch.cuda.stream(st):
        opt.zero_grad()
        logits = model(g, feat)
        loss = P.multi_loss_indexed(logits, lab, idx, w)
        loss.backward()
        opt.step()


models = [make(70 + i) for i in range(maxr)]
for r in range(1, maxr + 1):
    ms_ = models[:r]
    for _ in range(5):
        for m in ms_: epoch(m)
    torch.cuda.synchronize()
    steps = 40
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        for m in ms_: epoch(m)
    for m in ms_: torch.cuda.current_stream().wait_stream(m[3])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{r} concurrent models: {r * steps / (ms * 1e-3):.1f} epochs/s aggregate ({ms / steps:.3f} ms per round of {r} epochs), host wall {time.perf_counter() - t0:.3f} s")

# ---- the same with every model's epoch as one CUDA-graph replay (epoch.TrainStep) on its own stream: no host work per kernel ----
from plagnn_b200.epoch import TrainStep
steps_g = []
for i in range(maxr):
    torch.manual_seed(170 + i)
    model = P.GNN32(503, 400, 300, 200, 100, 12).to(dev)
    lab_rows = prob.labelled.copy(); rng.shuffle(lab_rows)
    idx = torch.as_tensor(np.sort(lab_rows[: len(lab_rows) * 9 // 10]), device=dev)
    steps_g.append((TrainStep(model, g, feat, lab, idx, w, lr=5e-5), torch.cuda.Stream(device=dev)))
for r in (1, 2, 4, 6, 8):
    if r > maxr:
        break
    ms_ = steps_g[:r]
    for _ in range(5):
        for ts, st in ms_:
            with torch.cuda.stream(st):
                ts.step()
    torch.cuda.synchronize()
    steps = 40
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _, st in ms_: st.wait_stream(torch.cuda.current_stream())
    s.record()
    for _, st in ms_: st.wait_stream(torch.cuda.current_stream())
    for _ in range(steps):
        for ts, st in ms_:
            with torch.cuda.stream(st):
                ts.step()
    for _, st in ms_: torch.cuda.current_stream().wait_stream(st)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{r} concurrent models as graph replays: {r * steps / (ms * 1e-3):.1f} epochs/s aggregate ({ms / steps:.3f} ms per round of {r} epochs)")
