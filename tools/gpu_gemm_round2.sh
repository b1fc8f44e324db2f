#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/gemm_debug.log; : > $L
for shape in "24041 503 503 0 0" "24064 512 8192 0 0" "400 503 24041 1 1"; do
  for d in 0 1 2 3; do PLAGNN_TMA_DEBUG=$d timeout 120 python tools/gemm_once.py $shape >> $L 2>&1; done
done
cat $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -c 3 -o gpurun_out/prof_tma_fwd python tools/gemm_once.py 24041 503 503 0 0 1 > gpurun_out/ncu_tma.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_tma.log
