#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
for CH in 1 4 8; do
  echo "== scaled full N=$N chunks=$CH"
  PLAGNN_DIST_CHUNKS=$CH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$CH bench.py --gpus $N --workload scaled --steps 8 --warmup 3 > gpurun_out/scaled_full_n${N}_c$CH.json 2> gpurun_out/s.err
  echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n${N}_c$CH.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), 'spmm ms', round(d['roofline']['spmm_ms_per_step'],3), 'edges/s', d['spmm_edges_per_s'])"; tail -2 gpurun_out/s.err | cut -c1-300
done
