#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"
tail -8 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py 2 > gpurun_out/gemm_bench_cg2.log 2>&1; echo "bench cg2 exit $?"; cat gpurun_out/gemm_bench_cg2.log
