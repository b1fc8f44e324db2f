#!/bin/bash
# round 2, seventh GPU call (2 GPUs): narrow kernel rewrite (tests + timing), peer-memory exchange (round trip + partitions), pipeline test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest_p2p.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_p2p.log
PLAGNN_SPMM_NARROW=1 timeout 400 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow2.json 2> gpurun_out/r2_narrow2.err; echo "narrow timing exit $?"; cut -c1-700 gpurun_out/r2_narrow2.json; tail -3 gpurun_out/r2_narrow2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/dist_sweep.py --steps 5 --ctas 0 --chunks 2 --feat 64 > gpurun_out/r2_sweep_n2_f64.jsonl 2> gpurun_out/r2_sweep_n2_f64.err
echo "sweep n2 (F=64: 32 columns per rank) exit $?"; tail -3 gpurun_out/r2_sweep_n2_f64.err
python - <<'PY'
import json
for line in open("gpurun_out/r2_sweep_n2_f64.jsonl"):
    if not line.startswith("{"): continue
    v=json.loads(line)
    if "mode" not in v: print(v); continue
    print(v["mode"], v["reducer"], v.get("exchange"), "ms %.2f nocomm %.2f exposed %.2f agg %.2f" % (v["ms_per_step"], v["ms_per_step_without_collectives"], v["exposed_exchange_ms"], v["aggregation_ms_per_step"]), v["collective_ms_per_step"], v["check"]["out_rel_err"], v["check"]["grad_rel_err_max"], [(k["kernel"], k["ms_per_step"]) for k in v["kernels"][:6]])
PY
