#!/bin/bash
mkdir -p gpurun_out
(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > gpurun_out/smi_ring.csv 2>&1 &) 
sleep 0.5
timeout 120 tools/_bin/mma_probe ring 2>&1 | tee gpurun_out/mma_ring.log
sleep 0.3
python - <<'PY'
import torch, time, os, sys
sys.path.insert(0, os.getcwd())
from plagnn_b200 import ops
dev = torch.device("cuda:0")
m, n, k = 24064, 512, 8192
a = ops.aligned(torch.randn(m, k, device=dev)); b = ops.aligned(torch.randn(n, k, device=dev)); out = ops.alloc(m, n, dev)
pairs = [(a, 0, b, 0, k)]
for _ in range(3): ops.gemm(m, n, pairs, out=out, backend=ops.GEMM_TMA)
torch.cuda.synchronize(); t0 = time.time()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(1500): ops.gemm(m, n, pairs, out=out, backend=ops.GEMM_TMA)
e.record(); torch.cuda.synchronize()
print(f"gemm loop: {s.elapsed_time(e)/1500:.4f} ms per GEMM, wall {time.time()-t0:.2f}s")
PY
sleep 0.3
pkill -x nvidia-smi
awk -F, 'NR>1{print $1","$3","$4}' gpurun_out/smi_ring.csv | sort | uniq -c | sort -rn | head -20
