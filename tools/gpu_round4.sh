#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 --cpu-seconds 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("gemm", d["gemm"]); print("spmm", d["spmm"]); print("cpu", d.get("cpu_baseline"))
for k in json.load(open("gpurun_out/bench_kernels_n1.json"))[:45]: print(k)
PY
tail -5 gpurun_out/bench.err
