"""GEMM backends on the PPI layer shapes: first-generation tcgen05 kernel (SIMT loader warps) vs the TMA-fed CTA-pair
kernel.  CUDA events, warm.
    python tools/gemm_bench.py [cg]  > gpurun_out/gemm_bench.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plagnn_b200 import ops

dev = torch.device("cuda:0")
if len(sys.argv) > 1:
    os.environ["PLAGNN_TMA_CG"] = sys.argv[1]


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


N = 24041
shapes = [(N, 503, 503, 0, 0, 1), (N, 400, 503, 0, 0, 2), (N, 400, 400, 0, 0, 1), (N, 300, 400, 0, 0, 2),
          (N, 503, 400, 0, 1, 1), (N, 400, 350, 0, 1, 2), (400, 503, N, 1, 1, 1), (503, 503, N, 1, 1, 1),
          (300, 400, N, 1, 1, 1), (200, 300, N, 1, 1, 1), (N, 100, 200, 0, 0, 1), (8192, 8192, 8192, 0, 0, 1)]
print(f"PLAGNN_TMA_CG={os.environ.get('PLAGNN_TMA_CG', '2')}")
for (m, n, k, at, bt, pr) in shapes:
    a = ops.aligned(torch.randn((k, m) if at else (m, k), device=dev))
    b = ops.aligned(torch.randn((k, n) if bt else (n, k), device=dev))
    out = ops.alloc(m, n, dev)
    fl = 2.0 * m * n * k * pr
    row = f"m={m:6d} n={n:5d} k={k:6d} at={at} bt={bt} pairs={pr}: "
    pairs = [(a, at, b, bt, k)] * pr
    ref = None
    for name, be in (("tcgen05", ops.GEMM_TCGEN05), ("tma", ops.GEMM_TMA)):
        ms = timeit(lambda: ops.gemm(m, n, pairs, out=out, backend=be))
        row += f"{name} {ms:.4f} ms {fl / ms / 1e9:.1f} TF | "
        if ref is None:
            ref = out.clone()
        else:
            row += f"(diff {((out - ref).abs().max() / ref.abs().max()).item():.1e}) "
    print(row, flush=True)
