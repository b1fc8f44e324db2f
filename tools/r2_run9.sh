#!/bin/bash
# round 2, ninth GPU call (1 GPU): restored narrow kernel (tests, U = 4 vs 8), Pearson kernel, preprocessing timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py tests/test_gpu_preprocess.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2_pytest_9.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_9.log
for u in 4 8; do
  PLAGNN_SPMM_NARROW_U=$u PLAGNN_TIME_NARROW_ONLY=1 timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow_u$u.json 2> gpurun_out/r2_narrow_u$u.err; echo "U=$u exit $?"; cat gpurun_out/r2_narrow_u$u.json
done
