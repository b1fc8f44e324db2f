#!/bin/bash
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'])"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload scaled --steps 6 --warmup 3 > gpurun_out/scaled_full_n2.json 2> gpurun_out/scaled_n2.err; echo "scaled n2 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), d['roofline'])"
timeout 600 python bench.py --workload scaled --steps 6 --warmup 3 > gpurun_out/scaled_full_n1.json 2> gpurun_out/scaled_n1.err; echo "scaled n1 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2), d['roofline']); print(d['kernels'])"
