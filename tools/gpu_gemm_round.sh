#!/bin/bash
# TMA GEMM bring-up: parity tests then timings, each under its own timeout
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k "tma or lo" > gpurun_out/pytest_tma.log 2>&1; echo "pytest tma exit $?"
tail -15 gpurun_out/pytest_tma.log
timeout 300 python tools/gemm_bench.py 2 > gpurun_out/gemm_bench_cg2.log 2>&1; echo "bench cg2 exit $?"; cat gpurun_out/gemm_bench_cg2.log
timeout 300 python tools/gemm_bench.py 1 > gpurun_out/gemm_bench_cg1.log 2>&1; echo "bench cg1 exit $?"; cat gpurun_out/gemm_bench_cg1.log
