#!/bin/bash
# round 2, fourth GPU call (1 GPU): epoch graph tests, loss tests, bench with the graph-replay leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_epoch.py tests/test_gpu_loss_adam.py tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r2_pytest_epoch.log 2>&1
echo "pytest epoch exit $?"; tail -25 gpurun_out/r2_pytest_epoch.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-partitioned > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench exit $?"; tail -3 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench4.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "graph", d["graph_replay"], "conc", d["concurrent_models"]["value"], "launches", d["gpu_launches"])
PY
