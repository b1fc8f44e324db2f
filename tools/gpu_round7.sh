#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_loss_adam.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_scoring.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/pytest_scoring.log
timeout 200 python tools/host_bound.py 2>&1 | head -3
tools/gpu_bench_only.sh 2>&1 | head -12
