#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/gemm_debug2.log; : > $L
for shape in "24041 503 503 0 0" "24041 400 503 0 0" "24041 503 400 0 1"; do
  for d in 0 3 4; do PLAGNN_TMA_DEBUG=$d timeout 120 python tools/gemm_once.py $shape >> $L 2>&1; done
done
cat $L
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k "tma or lo or ex" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"
tail -3 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py 2 > gpurun_out/gemm_bench_cg2.log 2>&1; echo "bench cg2 exit $?"; cat gpurun_out/gemm_bench_cg2.log
