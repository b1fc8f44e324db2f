#!/bin/bash
# round 2: the bench line (PPI replicas + partitioned configs[3] block) at N GPUs, as the driver launches it
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench n1 exit $?"
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2967$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err; echo "bench n$N exit $?"
fi
tail -3 gpurun_out/r2_final_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_final_bench_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "graph", d["graph_replay"].get("value"), "conc", d["concurrent_models"]["value"])
p=d["partitioned"]
print("winner", p["winner"], "checks_ok", p["checks_ok"], "p2p err", p.get("p2p_wait_gave_up_at_seq"), "nccl", p["nccl_alone"])
for v in p["variants"]:
    print("  ", v["mode"], v["reducer"], v.get("exchange"), "ms %.2f" % v["ms_per_step"], "single", v.get("single_gpu_ms_per_step_same_run"), "eff", v.get("strong_scaling_efficiency"), "exposed %.2f" % v["exposed_exchange_ms"], v["check"] and (v["check"]["out_rel_err"], v["check"]["grad_rel_err_max"], v["check"]["ok"]), v.get("aggregation_dram"))
PY
