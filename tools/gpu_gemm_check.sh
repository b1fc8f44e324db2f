#!/bin/bash
# GEMM parity + model parity, isolated GEMM timings, bench; optional per-CTA trace with the diagnostics build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gemm.log
timeout 300 python tools/gemm_bench.py 2 2>&1 | head -10
if [ -f pla-gnn_b200/libplagnn_diag.so ]; then
  for s in "24041 503 503 0 0" "400 503 24041 1 1"; do PLAGNN_LIB_PATH=$PWD/pla-gnn_b200/libplagnn_diag.so python tools/gemm_trace.py $s 2>&1 | head -5; done
fi
tools/gpu_bench_only.sh 2>&1 | head -9
