#!/bin/bash
# 8-GPU evidence: replicas of the PPI-shaped epoch (default bench) and the row-partitioned scaled graph
N=${1:-8}
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'], d['clocks'], d.get('concurrent_models'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --workload scaled --steps 8 --warmup 3 > gpurun_out/scaled_full_n$N.json 2> gpurun_out/scaled_n$N.err; echo "scaled n$N exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), 'edges/s', d.get('spmm_edges_per_s'), d['roofline']); print(d['kernels'])"
tail -2 gpurun_out/scaled_n$N.err | cut -c1-300
