#!/bin/bash
# round 2, last session, run 1 (1 GPU): the whole GPU suite on the new defaults (pipelined narrow aggregation, warm-up warp of
# the TMA GEMM), then the aggregation variants on the 100 M-edge graph, the GEMM warm-up A/B, and the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2_16_pytest.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/r2_16_pytest.log
timeout 600 python tools/spmm_variants_time.py > gpurun_out/r2_spmm_variants.json 2> gpurun_out/r2_spmm_variants.err; echo "variants exit $?"; tail -60 gpurun_out/r2_spmm_variants.err
timeout 400 python tools/gemm_warm_ab.py > gpurun_out/r2_gemm_warm_ab.json 2> gpurun_out/r2_gemm_warm_ab.err; echo "warm ab exit $?"; tail -20 gpurun_out/r2_gemm_warm_ab.err
for warm in 1 0; do
  PLAGNN_TMA_WARM=$warm timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline $([ $warm = 0 ] && echo "--no-partitioned --no-pipeline") > gpurun_out/r2_16_bench_warm$warm.json 2> gpurun_out/r2_16_bench_warm$warm.err; echo "bench warm=$warm exit $?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_16_bench_warm$warm.json").read().strip().splitlines() if l.startswith("{")][-1])
print("warm=$warm value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "graph", d["graph_replay"].get("value"), "conc", d["concurrent_models"]["value"], "gemm", d["gemm"]["ms_per_step"], "spmm", d["spmm"], "roofline", d["roofline"]["frac"])
p=d.get("partitioned")
for v in (p or {}).get("variants", []):
    print("  ", v["mode"], v["reducer"], "ms %.2f" % v["ms_per_step"], "agg", v.get("aggregation_ms_per_step"))
PY
done
