"""Accuracy of the TMA GEMM vs float64 on layer-like shapes and data (norm-wise max relative error).
    PLAGNN_TMA_SINGLE_ACC=1 python tools/gemm_accuracy.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plagnn_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
N = 24041
print("single_acc =", os.environ.get("PLAGNN_TMA_SINGLE_ACC", "0"))
def run(m, n, k, at, bt, kind):
    if kind == "randn":
        a = torch.randn((k, m) if at else (m, k), device=dev); b = torch.randn((k, n) if bt else (n, k), device=dev)
    elif kind == "relu":     # post-ReLU activations (non-negative: no cancellation, biased sums) x small weights
        a = torch.relu(torch.randn((k, m) if at else (m, k), device=dev)) + 0.1; b = torch.rand((k, n) if bt else (n, k), device=dev) * 0.05
    else:                    # all positive, large dynamic range
        a = torch.rand((k, m) if at else (m, k), device=dev) * 8 + 1; b = torch.rand((k, n) if bt else (n, k), device=dev) + 0.5
    a, b = ops.aligned(a), ops.aligned(b)
    got = ops.gemm(m, n, [(a, at, b, bt, k)], backend=ops.GEMM_TMA)
    A = (a[:, :m].double().t() if at else a[:, :k].double()); B = (b[:, :n].double() if bt else b[:, :k].double().t())
    ref = A @ B
    err = ((got[:, :n].double() - ref).abs().max() / ref.abs().max()).item()
    print(f"m={m:6d} n={n:4d} k={k:6d} at={at} bt={bt} {kind:6s}: max rel err {err:.2e}")
for kind in ("randn", "relu", "pos"):
    run(4096, 503, 503, 0, 0, kind)
    run(4096, 400, 1006, 0, 0, kind)
    run(4096, 512, 1280, 0, 0, kind)
    run(400, 503, N, 1, 1, kind)
    run(503, 503, N, 1, 1, kind)
