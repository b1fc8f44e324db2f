#!/bin/bash
# per-CTA timelines of the TMA GEMM (diagnostics build) on the products of the epoch, incl. the second tile of a pair
mkdir -p gpurun_out
export PLAGNN_LIB_PATH=pla-gnn_b200/libplagnn_diag.so
for shape in "24041 503 503 0 0" "24041 400 400 0 0" "24041 503 400 0 1" "400 503 24041 1 1" "300 400 24041 1 1" "24041 100 200 0 0"; do
  timeout 120 python tools/gemm_trace.py $shape 2>&1 | head -12
done > gpurun_out/r2_gemm_trace.log 2>&1
cat gpurun_out/r2_gemm_trace.log
