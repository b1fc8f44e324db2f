#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py 2 2>&1 | head -10
tools/gpu_bench_only.sh 2>&1 | head -9
