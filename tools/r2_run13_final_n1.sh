#!/bin/bash
# round 2: the whole GPU suite, smoke, then the default bench line on one GPU (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_final_ref.json 2>&1; echo "reference arm exit $?"; cut -c1-200 gpurun_out/r2_final_ref.json
tools/r2_run12_final_multi.sh 1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("roofline", d["roofline"]["frac"], "agg frac_dram", d["roofline_aggregation"]["frac_dram"], "frac_l2", d["roofline_aggregation"]["frac_l2"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("config2", d["config2_pipeline"]); print("cpu", d.get("cpu_baseline"))
print("kernels", [(k["kernel"], k["avg_ms"]) for k in d["kernels"][:6]], d["gemm"]["ms_per_step"], d["spmm"])
PY
