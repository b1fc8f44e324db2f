#!/bin/bash
# round 2, 8-GPU call #2: partition sweep with the peer-memory exchange and the per-launch-shape aggregation times
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/dist_sweep.py --steps 5 --ctas 0 --chunks 2 > gpurun_out/r2_sweep2_n8.jsonl 2> gpurun_out/r2_sweep2_n8.err
echo "sweep exit $?"; tail -3 gpurun_out/r2_sweep2_n8.err; wc -l gpurun_out/r2_sweep2_n8.jsonl
