#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/gemm_trace4.log; : > $L
for shape in "24041 503 503 0 0" "400 503 24041 1 1"; do
  PLAGNN_TMA_DEBUG=5 timeout 60 python tools/gemm_trace.py $shape 2>&1 | head -12 >> $L
  timeout 60 python tools/gemm_trace.py $shape 2>&1 | head -12 >> $L
done
cat $L
