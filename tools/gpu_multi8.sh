#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
echo "== replicas N=$N"
run 29521 --steps 30 --warmup 8 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'])"; tail -2 gpurun_out/bench_n$N.err
echo "== scaled full N=$N"
NCCL_DEBUG=WARN run 29522 --workload scaled --steps 8 --warmup 3 > gpurun_out/scaled_full_n$N.json 2> gpurun_out/scaled_full_n$N.err
echo "exit $?"; tail -1 gpurun_out/scaled_full_n$N.json | cut -c1-2200; tail -2 gpurun_out/scaled_full_n$N.err
