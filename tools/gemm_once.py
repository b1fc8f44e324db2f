"""One GEMM shape through plagnn_gemm_ex, timed (CUDA events) — for PLAGNN_TMA_DEBUG experiments and ncu captures.
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
al, bl = ops.tf32_lo(a), ops.tf32_lo(b)
out, out_lo = ops.alloc(m, n, dev), ops.alloc(m, n, dev)
ex = [(a, al, at, b, bl, bt, k)]
for _ in range(3):
    ops.gemm_ex(m, n, ex, out=out, out_lo=out_lo)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    ops.gemm_ex(m, n, ex, out=out, out_lo=out_lo)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
print(f"debug={os.environ.get('PLAGNN_TMA_DEBUG', '0')} cg={os.environ.get('PLAGNN_TMA_CG', '2')} m={m} n={n} k={k} at={at} bt={bt}: "
      f"{ms:.4f} ms {2.0 * m * n * k / ms / 1e9:.1f} TF")
