#!/bin/bash
# full GPU suite with programmatic dependent launch on (default), then the bench with it on and off
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for pdl in 1 0 1 0; do
PLAGNN_PDL=$pdl timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_pdl$pdl.json").read().strip().splitlines()[-1])
print("pdl=$pdl", {k: d[k] for k in ("value", "ms_per_step", "ms_per_step_profiled")}, "e2e ms", d["e2e"]["ms_per_step"], d.get("concurrent_models"))
PY
done
