#!/bin/bash
# round 2, 8-GPU call: partition sweep, then the default bench line at N = 8 (partitioned block included)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/dist_sweep.py --steps 5 > gpurun_out/r2_sweep_n8.jsonl 2> gpurun_out/r2_sweep_n8.err
echo "sweep exit $?"; tail -3 gpurun_out/r2_sweep_n8.err; wc -l gpurun_out/r2_sweep_n8.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n8 exit $?"; tail -3 gpurun_out/r2_bench_n8.err; cut -c1-300 gpurun_out/r2_bench_n8.json
