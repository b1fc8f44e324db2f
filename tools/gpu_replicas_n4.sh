#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench n4 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','ms_per_step_profiled','n_gpus','scaling')}, d['e2e'], d['clocks'], d.get('concurrent_models'))"
