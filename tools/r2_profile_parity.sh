#!/bin/bash
# ncu evidence for the default (parity) GEMM mode: launch list of bench.py, --set full of the two-chain kernel
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned --no-pipeline > gpurun_out/r2p_plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 900 --csv --log-file gpurun_out/launches_r2_parity.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned --no-pipeline > gpurun_out/r2p_ncu_launches.log 2>&1; echo "ncu launches exit $?"
python tools/summarize_ncu.py launches gpurun_out/launches_r2_parity.csv > gpurun_out/launches_r2_parity.md 2>&1
timeout 300 python tools/kernels_once.py gemm > gpurun_out/r2p_k1.log 2>&1; echo "kernels_once exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"gemm_tma" -c 8 -f -o gpurun_out/prof_r2_parity python tools/kernels_once.py gemm > gpurun_out/r2p_ncu.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/prof_r2_parity.ncu-rep --page raw --csv > gpurun_out/prof_r2_parity.raw.csv 2> /dev/null
python tools/summarize_ncu.py full gpurun_out/prof_r2_parity.ncu-rep > gpurun_out/prof_r2_parity.md 2>&1
rm -f gpurun_out/prof_r2_parity.ncu-rep
cat gpurun_out/launches_r2_parity.md | head -14; cat gpurun_out/prof_r2_parity.md
