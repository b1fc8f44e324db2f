#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
for i in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','ms_per_step_profiled','n_gpus','scaling')}, d['e2e'], d['clocks'])"
done
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1b.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1b.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','ms_per_step_profiled','n_gpus','scaling')}, d['e2e'], d['clocks'])"
nproc; python -c "import os; print(os.sched_getaffinity(0))"
