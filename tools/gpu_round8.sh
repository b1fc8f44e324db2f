#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py -x -q > gpurun_out/pytest_scoring.log 2>&1; echo "pytest scoring exit $?"
tail -3 gpurun_out/pytest_scoring.log
timeout 60 python tools/gemm_trace.py 24041 503 503 0 0 2>&1 | head -9
