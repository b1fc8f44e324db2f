#!/bin/bash
# N GPUs: feature partition with the layers taken in pairs (PLAGNN_DIST_COLS_ALT) on / off, in-run numeric check included
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2_24_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2_24_pytest.log
fi
for alt in 1 0; do
  PLAGNN_DIST_COLS_ALT=$alt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$alt tools/dist_sweep.py --steps 5 --ctas 0 --chunks 2 --modes cols --reducers sum --exchange p2p > gpurun_out/r2_24_sweep_n${N}_alt$alt.jsonl 2> gpurun_out/r2_24_sweep_n${N}_alt$alt.err; echo "sweep alt=$alt exit $?"
  python - <<PY
import json
for l in open("gpurun_out/r2_24_sweep_n${N}_alt$alt.jsonl"):
    l=l.strip()
    if not l.startswith("{"): continue
    v=json.loads(l)
    print("alt=$alt", v.get("mode"), v.get("reducer"), v.get("exchange"), "ms", v.get("ms_per_step"), "no-exchange", v.get("ms_per_step_without_collectives"), "exposed", v.get("exposed_exchange_ms"), "agg", v.get("aggregation_ms_per_step"), v.get("check"))
PY
  tail -3 gpurun_out/r2_24_sweep_n${N}_alt$alt.err
done
