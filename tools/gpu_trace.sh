#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/gemm_trace2.log; : > $L
for shape in "24041 503 503 0 0" "24064 512 512 0 0"; do
  PLAGNN_TMA_COMPANION=1 timeout 60 python tools/gemm_trace.py $shape 2>&1 | head -6 >> $L
done
cat $L
