#!/bin/bash
# ncu evidence for profiles/ on the final code of the round: launch list of bench.py, --set full of the hot kernels, bench
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 700 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu launches exit $?"
timeout 300 python tools/kernels_once.py all > gpurun_out/k1.log 2>&1; echo "kernels_once exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma_kernel|spmm_kernel|spmm_max_scatter|narrow_" -o gpurun_out/prof_r1f python tools/kernels_once.py all > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
tools/gpu_bench_only.sh 2>&1 | cut -c1-700
