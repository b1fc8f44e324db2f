#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 10 --cpu-seconds 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "ms_per_step_profiled", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("gemm", d["gemm"]); print("spmm", d["spmm"]); print("cpu", d.get("cpu_baseline"))
ks = json.load(open("gpurun_out/bench_kernels_n1_s50.json"))
tot = {}
for k in ks:
    tot[k["kernel"].split("/")[0]] = tot.get(k["kernel"].split("/")[0], 0) + k["avg_ms"] * k["calls_per_step"]
print({k: round(v, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
for k in ks[:40]: print(k["kernel"], k["calls_per_step"], k["avg_ms"])
PY
tail -5 gpurun_out/bench.err
