#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_model.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
timeout 200 python tools/spmm_time.py
for v in 1; do PLAGNN_LIB_PATH=$PWD/pla-gnn_b200/libplagnn_v$v.so timeout 200 python tools/spmm_time.py; done
tools/gpu_bench_only.sh 2>&1 | sed -n 1,9p | cut -c1-400
