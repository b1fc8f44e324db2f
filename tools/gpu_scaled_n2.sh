#!/bin/bash
# row-partitioned scaled graph on 2 GPUs (NCCL exchange interleaved with the chained kernels)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --workload scaled --steps 6 --warmup 3 > gpurun_out/scaled_full_n2.json 2> gpurun_out/scaled_n2.err; echo "scaled n2 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), 'spmm ms', round(d['roofline']['spmm_ms_per_step'],3)); print(d.get('parity'), d.get('loss'))"
tail -2 gpurun_out/scaled_n2.err | cut -c1-300
