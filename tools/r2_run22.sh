#!/bin/bash
# double-buffered GEMM by the cost model: A/B per product, whole GPU suite, bench
mkdir -p gpurun_out
timeout 300 python tools/gemm_db_ab.py > gpurun_out/r2_gemm_db_ab.json 2> gpurun_out/r2_gemm_db_ab.err; echo "ab exit $?"; tail -16 gpurun_out/r2_gemm_db_ab.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_22_pytest.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2_22_pytest.log
for db in 1 0; do
  PLAGNN_TMA_DB=$db timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline $([ $db = 0 ] && echo "--no-partitioned --no-pipeline") > gpurun_out/r2_22_bench_db$db.json 2> gpurun_out/r2_22_bench_db$db.err; echo "bench db=$db exit $?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_22_bench_db$db.json").read().strip().splitlines() if l.startswith("{")][-1])
print("db=$db value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "graph", d["graph_replay"].get("value"), "conc", d["concurrent_models"]["value"], "gemm", d["gemm"]["ms_per_step"], "spmm", d["spmm"], "roofline", d["roofline"]["frac"])
p=d.get("partitioned")
for v in (p or {}).get("variants", []):
    print("  ", v["mode"], v["reducer"], "ms %.2f" % v["ms_per_step"], "agg", v.get("aggregation_ms_per_step"))
PY
done
