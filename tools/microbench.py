"""Per-kernel timings (CUDA events, warm, repeated) for the PPI-shaped layer shapes, both GEMM backends,
plus host-side enqueue cost of one epoch.  python tools/microbench.py > gpurun_out/microbench.log"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import plagnn_b200 as P
from plagnn_b200 import ops, synth

dev = torch.device("cuda:0")


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def gemm_case(m, n, k, at, bt, backend, pairs=1):
    a = ops.aligned(torch.randn((k, m) if at else (m, k), device=dev))
    b = ops.aligned(torch.randn((k, n) if bt else (n, k), device=dev))
    pr = [(a, at, b, bt, k)] * pairs
    out = ops.alloc(m, n, dev)
    ms = timeit(lambda: ops.gemm(m, n, pr, out=out, backend=backend))
    fl = 2.0 * m * n * k * pairs
    return ms, fl / ms / 1e9


if __name__ == "__main__":
    N = 24041
    print("== GEMM shapes (ms, TFLOP/s fp32-equivalent)")
    shapes = [(N, 503, 503, 0, 0, 1), (N, 400, 503, 0, 0, 2), (N, 400, 400, 0, 0, 1), (N, 300, 400, 0, 0, 2),
              (N, 503, 400, 0, 0, 1), (N, 400, 300, 0, 0, 1), (400, 503, N, 1, 1, 1), (503, 503, N, 1, 1, 1),
              (300, 400, N, 1, 1, 1), (N, 100, 200, 0, 0, 1), (8192, 8192, 8192, 0, 0, 1)]
    for (m, n, k, at, bt, pr) in shapes:
        row = f"m={m:6d} n={n:5d} k={k:6d} at={at} bt={bt} pairs={pr}: "
        for name, be in (("simt", ops.GEMM_SIMT), ("tcgen05", ops.GEMM_TCGEN05)):
            ms, tf = gemm_case(m, n, k, at, bt, be, pr)
            row += f"{name} {ms:.4f} ms {tf:.1f} TF | "
        print(row, flush=True)

    print("== SpMM on the PPI-shaped graph")
    prob = synth.ppi_problem(state="inter")
    g = P.graph((prob.ppi_row, prob.ppi_col), num_nodes=prob.num_nodes).add_self_loop().to(dev)
    csc, csr = g.csc(), g.csr()
    ep = csc.num_edges
    print("items/hubs/slots", csc.counts, "max in-degree", int(csc.degrees.max()))
    for f in (503, 400, 300, 256, 128, 64):
        x = ops.aligned(torch.relu(torch.randn(N, f, device=dev)))
        ms = timeit(lambda: ops.spmm_max_fwd(csc, x))
        out, arg = ops.spmm_max_fwd(csc, x)
        dz = ops.aligned(torch.randn(N, f, device=dev))
        ms_b = timeit(lambda: ops.spmm_max_bwd(dz, arg, out, N))
        ms_g = timeit(lambda: ops.spmm_max_bwd_gather(csr, dz, arg, out))
        ms_s = timeit(lambda: ops.spmm_sum(csc, x))
        alg = 4 * f * ep + 8 * f * N + 4 * ep + 4 * (N + 1)
        print(f"F={f}: max_fwd {ms:.4f} ms ({alg / ms / 1e6:.0f} GB/s alg, {ep / ms / 1e6:.2f} Gedge/s) | bwd scatter {ms_b:.4f} | "
              f"bwd gather {ms_g:.4f} | sum {ms_s:.4f} ms ({(alg - 4 * f * N) / ms_s / 1e6:.0f} GB/s alg)", flush=True)

    print("== one epoch: host enqueue time vs device time")
    ids = list(range(N))
    gg = P.create_graph(prob.scipy_ppi(), prob.ecc, prob.gcn, prob.scipy_loc(), prob.expr, ids).to(dev)
    model = P.GNN32(503, 400, 300, 200, 100, 12).to(dev)
    opt = P.FusedAdam(model.parameters(), lr=5e-5)
    w = P.weight_cal(prob.loc)
    idx = torch.as_tensor(prob.labelled[::2], device=dev)

    def epoch():
        opt.zero_grad()
        logits = model(gg, gg.ndata["feat"])
        loss = P.multi_loss_indexed(logits, gg.ndata["loc"], idx, w)
        loss.backward()
        opt.step()

    for be in ("simt", "tcgen05"):
        orig = ops.gemm
        forced = ops.GEMM_SIMT if be == "simt" else ops.GEMM_AUTO
        ops.gemm = lambda *a, **k: orig(*a, **{**k, "backend": forced})
        for _ in range(3):
            epoch()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            epoch()
        t_enq = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / 10
        print(f"{be}: host enqueue {t_enq * 1e3:.2f} ms/epoch, wall incl. drain {t_all * 1e3:.2f} ms/epoch, "
              f"device (events) {timeit(epoch, 10, 0):.2f} ms/epoch", flush=True)
        ops.gemm = orig
