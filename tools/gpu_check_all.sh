#!/bin/bash
# full GPU suite + smoke + bench (the round-end sequence of the driver)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
tools/gpu_bench_only.sh 2>&1 | head -10
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('concurrent', d.get('concurrent_models'))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
