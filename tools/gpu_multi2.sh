#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
for N in 1 2; do
 for CH in 1 4; do
  echo "== scaled full N=$N chunks=$CH"
  if [ $N = 1 ]; then PLAGNN_DIST_CHUNKS=$CH timeout 900 python bench.py --workload scaled --steps 6 --warmup 3 > gpurun_out/s.json 2> gpurun_out/s.err
  else PLAGNN_DIST_CHUNKS=$CH timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$CH bench.py --gpus $N --workload scaled --steps 6 --warmup 3 > gpurun_out/s.json 2> gpurun_out/s.err; fi
  echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/s.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), 'spmm ms', round(d['roofline']['spmm_ms_per_step'],3), 'edges/s', d['spmm_edges_per_s'])"; tail -2 gpurun_out/s.err | cut -c1-300
 done
done
