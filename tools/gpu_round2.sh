#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05 diagnostic"
timeout 300 python tools/diag_tcgen05.py > gpurun_out/diag_tcgen05.log 2>&1
echo "exit $?"; grep "rel err" gpurun_out/diag_tcgen05.log | awk '{print $NF}' | tr '\n' ' '; echo
echo "== tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?"; tail -6 gpurun_out/pytest_gpu.log
cat gpurun_out/grad_parity_tcgen05.txt
echo "== microbench"
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; cat gpurun_out/microbench.log
echo "== bench"
timeout 600 python bench.py --steps 30 --warmup 10 --cpu-seconds 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("gemm", d["gemm"]); print("spmm", d["spmm"]); print("cpu", d.get("cpu_baseline"))
PY
tail -5 gpurun_out/bench.err
