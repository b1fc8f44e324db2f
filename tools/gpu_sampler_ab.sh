#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --concurrent-models 1 > gpurun_out/b_ab.json 2> gpurun_out/b_ab.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/b_ab.json").read().strip().splitlines()[-1])
print("$1", round(d["ms_per_step"],4), round(d["ms_per_step_profiled"],4), round(d["e2e"]["ms_per_step"],4), d["clocks"])
PY
}
for rep in 1 2; do
PLAGNN_CLOCK_POLL_MS=20 run nvml20
PLAGNN_CLOCK_POLL_MS=60 run nvml60
PLAGNN_CLOCK_SOURCE=smi run smi
PLAGNN_CLOCK_POLL_MS=100000 run nvml_none
done
