#!/bin/bash
# round 2 (1 GPU): narrow kernel with the predicate-free inner loop — tests, then timing at 32 / 64 columns
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest_14.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_pytest_14.log
PLAGNN_TIME_NARROW_ONLY=1 timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow4.json 2> gpurun_out/r2_narrow4.err; echo "timing exit $?"; cat gpurun_out/r2_narrow4.json
