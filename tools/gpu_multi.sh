#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
echo "== replicas N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 8 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"; tail -3 gpurun_out/bench_n2.err
echo "== scaled small N=1"
timeout 600 python bench.py --workload scaled --nodes 200000 --edges 20000000 --steps 5 --warmup 3 > gpurun_out/scaled_small_n1.json 2> gpurun_out/scaled_small_n1.err
echo "exit $?"; cat gpurun_out/scaled_small_n1.json | cut -c1-1500; tail -3 gpurun_out/scaled_small_n1.err
echo "== scaled small N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload scaled --nodes 200000 --edges 20000000 --steps 5 --warmup 3 > gpurun_out/scaled_small_n2.json 2> gpurun_out/scaled_small_n2.err
echo "exit $?"; cat gpurun_out/scaled_small_n2.json | cut -c1-1500; tail -3 gpurun_out/scaled_small_n2.err
echo "== scaled full N=1"
timeout 900 python bench.py --workload scaled --steps 5 --warmup 3 > gpurun_out/scaled_full_n1.json 2> gpurun_out/scaled_full_n1.err
echo "exit $?"; cat gpurun_out/scaled_full_n1.json | cut -c1-1800; tail -3 gpurun_out/scaled_full_n1.err
echo "== scaled full N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload scaled --steps 5 --warmup 3 > gpurun_out/scaled_full_n2.json 2> gpurun_out/scaled_full_n2.err
echo "exit $?"; cat gpurun_out/scaled_full_n2.json | cut -c1-1800; tail -3 gpurun_out/scaled_full_n2.err
