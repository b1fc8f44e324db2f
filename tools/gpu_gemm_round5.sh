#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"
tail -4 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py 2 > gpurun_out/gemm_bench_cg2.log 2>&1; echo "bench cg2 exit $?"; cat gpurun_out/gemm_bench_cg2.log
L=gpurun_out/gemm_trace3.log; : > $L
for shape in "24041 503 503 0 0" "400 503 24041 1 1"; do
  timeout 60 python tools/gemm_trace.py $shape 2>&1 | head -7 >> $L
done
cat $L
