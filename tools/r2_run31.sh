#!/bin/bash
# 8 GPUs: feature partition, aggregation + peer-memory exchange fused block by block (PLAGNN_DIST_OVERLAP) on / off
N=${1:-8}
mkdir -p gpurun_out
for ov in 1 0; do
  PLAGNN_DIST_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$ov tools/dist_sweep.py --steps 5 --ctas 0 --chunks 2 --modes cols --reducers sum,max --exchange p2p > gpurun_out/r2_31_sweep_n${N}_ov$ov.jsonl 2> gpurun_out/r2_31_sweep_n${N}_ov$ov.err; echo "sweep overlap=$ov exit $?"
  python - <<PY
import json
for l in open("gpurun_out/r2_31_sweep_n${N}_ov$ov.jsonl"):
    l=l.strip()
    if not l.startswith("{"): continue
    v=json.loads(l)
    if v.get("mode") is None: continue
    print("overlap=$ov", v.get("mode"), v.get("reducer"), v.get("exchange"), "ms", round(v.get("ms_per_step"),2), "no-exchange", round(v.get("ms_per_step_without_collectives") or 0,2), "exposed", round(v.get("exposed_exchange_ms") or 0,2), "agg", round(v.get("aggregation_ms_per_step") or 0,2), (v.get("check") or {}).get("ok"))
PY
done
