#!/bin/bash
# ncu evidence for profiles/: launch list of bench.py and --set full of the hot kernels (each after its plain run exited 0)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 700 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu launches exit $?"
timeout 300 python tools/kernels_once.py all > gpurun_out/k1.log 2>&1; echo "kernels_once exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma_kernel|spmm_kernel|spmm_max_scatter|spmm_combine" -o gpurun_out/prof_r1e python tools/kernels_once.py all > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err; echo "sweep exit $?"; tail -3 gpurun_out/sweep.err
for s in "24041 503 503 0 0" "24041 400 1006 0 0" "24041 400 700 0 1" "400 503 24041 1 1"; do PLAGNN_LIB_PATH=$PWD/pla-gnn_b200/libplagnn_diag.so python tools/gemm_trace.py $s 2>&1 | head -5; done > gpurun_out/gemm_trace_final.log
timeout 200 python tools/gemm_bench.py 2 > gpurun_out/gemm_bench_final.log 2>&1
timeout 100 python tools/gemm_accuracy.py > gpurun_out/gemm_acc_final.log 2>&1
