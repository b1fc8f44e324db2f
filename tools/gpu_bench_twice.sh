#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$i.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_profiled")}, d["e2e"]["ms_per_step"], d.get("concurrent_models"))
PY
done
