#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_model.py tests/test_gpu_dist.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
timeout 600 python bench.py --workload scaled --steps 6 --warmup 3 > gpurun_out/scaled_full_n1.json 2> gpurun_out/scaled_n1.err; echo "scaled n1 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2)); print(d['kernels'][:3])"
tools/gpu_bench_only.sh 2>&1 | sed -n 1,3p | cut -c1-500
