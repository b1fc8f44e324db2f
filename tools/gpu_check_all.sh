#!/bin/bash
# full GPU suite + smoke + bench + scaled workload (N = 1) + reference arm: the round-end sequence of the driver and more
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
tools/gpu_bench_only.sh 2>&1 | head -4 | cut -c1-400
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('concurrent', d.get('concurrent_models'))"
timeout 600 python bench.py --workload scaled --steps 6 --warmup 3 > gpurun_out/scaled_full_n1.json 2> gpurun_out/scaled_n1.err; echo "scaled n1 exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/scaled_full_n1.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'epochs/s', round(d['value'],2)); print(d['kernels'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
