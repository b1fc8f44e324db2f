import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plagnn_b200 import ops
from tools.microbench import gemm_case
N = 24041
for (m, n, k, at, bt, pr) in [(N, 503, 503, 0, 0, 1), (N, 400, 503, 0, 0, 2), (400, 503, N, 1, 1, 1), (8192, 8192, 8192, 0, 0, 1)]:
    ms, tf = gemm_case(m, n, k, at, bt, ops.GEMM_TCGEN05, pr)
    print(f"dbg={os.environ.get('PLAGNN_TC_DEBUG','0')} m={m} n={n} k={k} at={at} pairs={pr}: {ms:.4f} ms {tf:.1f} TF", flush=True)
