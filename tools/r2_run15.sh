#!/bin/bash
# round 2 (1 GPU): lean loop of the wide sum kernel — tests, then timing with the lean loop on / off; narrow kernel U = 8 again
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k "not full_size" > gpurun_out/r2_pytest_15.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_pytest_15.log
for lean in 1 0; do
  PLAGNN_SPMM_SUM_LEAN=$lean timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_sum_lean$lean.json 2> gpurun_out/r2_sum_lean$lean.err; echo "lean=$lean exit $?"; cat gpurun_out/r2_sum_lean$lean.json
done
PLAGNN_SPMM_NARROW_U=8 PLAGNN_TIME_NARROW_ONLY=1 timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow5_u8.json 2>/dev/null; echo "U=8 exit $?"; cat gpurun_out/r2_narrow5_u8.json
