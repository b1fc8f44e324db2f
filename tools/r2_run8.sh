#!/bin/bash
# round 2, eighth GPU call (1 GPU): staged narrow kernel (tests + timing), unchanged-driver integration test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_spmm.py tests/test_gpu_integration.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2_pytest_stage.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/r2_pytest_stage.log
PLAGNN_DIST_SLAB_MB=0 timeout 400 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow3.json 2> gpurun_out/r2_narrow3.err; echo "narrow timing exit $?"; cut -c1-1200 gpurun_out/r2_narrow3.json; tail -3 gpurun_out/r2_narrow3.err
