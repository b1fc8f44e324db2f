#!/bin/bash
# round 2, sixth GPU call (1 GPU): slab passes (tests + timing at full scale), pipeline test, bench with all legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest_slab.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_slab.log
timeout 400 python tools/spmm_narrow_time.py > gpurun_out/r2_slab_time.json 2> gpurun_out/r2_slab_time.err; echo "slab timing exit $?"; cat gpurun_out/r2_slab_time.json; tail -3 gpurun_out/r2_slab_time.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench exit $?"; tail -3 gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench6.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "graph", d["graph_replay"]["value"], "conc", d["concurrent_models"]["value"])
print("config2", d["config2_pipeline"])
print("roofline", d["roofline"]["frac"], "agg", d["roofline_aggregation"]["frac_dram"], d["roofline_aggregation"]["frac_l2"])
print("cpu", d.get("cpu_baseline"))
PY
