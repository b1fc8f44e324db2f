"""One GEMM shape through the TMA-fed backend, timed (CUDA events) — for PLAGNN_TMA_DEBUG experiments and ncu captures.
    python tools/gemm_once.py m n k at bt [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plagnn_b200 import ops

dev = torch.device("cuda:0")
m, n, k, at, bt = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 20
a = ops.aligned(torch.randn((k, m) if at else (m, k), device=dev))
b = ops.aligned(torch.randn((k, n) if bt else (n, k), device=dev))
out = ops.alloc(m, n, dev)
pairs = [(a, at, b, bt, k)]
for _ in range(3):
    ops.gemm(m, n, pairs, out=out, backend=ops.GEMM_TMA)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    ops.gemm(m, n, pairs, out=out, backend=ops.GEMM_TMA)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
print(f"debug={os.environ.get('PLAGNN_TMA_DEBUG', '0')} cg={os.environ.get('PLAGNN_TMA_CG', '2')} m={m} n={n} k={k} at={at} bt={bt}: "
      f"{ms:.4f} ms {2.0 * m * n * k / ms / 1e9:.1f} TF")
