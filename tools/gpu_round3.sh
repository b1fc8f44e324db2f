#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_tcgen05.py > gpurun_out/diag_tcgen05.log 2>&1
echo "diag exit $?"; grep "rel err" gpurun_out/diag_tcgen05.log | awk '{print $NF}' | tr '\n' ' '; echo
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; grep -E "tcgen05|epoch" gpurun_out/microbench.log
