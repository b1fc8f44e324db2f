#!/bin/bash
# GPU suite + the default bench (no scaled workload, no reference arm): the quick check after a kernel change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
tools/gpu_bench_only.sh 2>&1 | cut -c1-600
