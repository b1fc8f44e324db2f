#!/bin/bash
# round 2, 2-GPU call: aggregation + peer-memory exchange fused block by block (tests, then overlap on / off at 32 columns per rank)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest_11.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/r2_pytest_11.log
for ov in 1 0; do
PLAGNN_DIST_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$ov tools/dist_sweep.py --steps 5 --ctas 0 --chunks 2 --feat 64 --modes cols --exchange p2p > gpurun_out/r2_sweep_n2_ov$ov.jsonl 2> gpurun_out/r2_sweep_n2_ov$ov.err
echo "sweep overlap=$ov exit $?"; tail -2 gpurun_out/r2_sweep_n2_ov$ov.err
python - <<PY
import json
for line in open("gpurun_out/r2_sweep_n2_ov$ov.jsonl"):
    if not line.startswith("{"): continue
    v=json.loads(line)
    if "mode" not in v: print(v); continue
    print(v["mode"], v["reducer"], v.get("exchange"), "ms %.2f nocomm %.2f exposed %.2f agg %.2f" % (v["ms_per_step"], v["ms_per_step_without_collectives"], v["exposed_exchange_ms"], v["aggregation_ms_per_step"]), v["collective_ms_per_step"], v["check"]["out_rel_err"], v["check"]["grad_rel_err_max"], v["check"]["ok"])
PY
done
