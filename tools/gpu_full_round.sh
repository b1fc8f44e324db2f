#!/bin/bash
# full validation round: GPU tests, smoke, bench, ncu launch list, ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 --cpu-seconds 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_profiled", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("gemm", d["gemm"]); print("spmm", d["spmm"]); print("cpu", d.get("cpu_baseline"))
for k in json.load(open("gpurun_out/bench_kernels_n1.json"))[:30]: print(k)
PY
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu launches exit $?"
timeout 300 python tools/kernels_once.py all > gpurun_out/k1.log 2>&1; echo "kernels_once exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma_kernel|spmm_kernel|spmm_max_scatter|spmm_combine" -o gpurun_out/prof_r1b python tools/kernels_once.py all > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
