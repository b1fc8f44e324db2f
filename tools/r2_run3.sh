#!/bin/bash
# round 2, third GPU call (1 GPU): narrow-row aggregation kernel (tests + A/B timing), full-size parity tests, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_dist.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest_spmm.log 2>&1
echo "pytest spmm+dist exit $?"; tail -8 gpurun_out/r2_pytest_spmm.log
PLAGNN_SPMM_NARROW=0 timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow0.json 2> gpurun_out/r2_narrow0.err; echo "narrow=0 exit $?"; cat gpurun_out/r2_narrow0.json
PLAGNN_SPMM_NARROW=1 timeout 300 python tools/spmm_narrow_time.py > gpurun_out/r2_narrow1.json 2> gpurun_out/r2_narrow1.err; echo "narrow=1 exit $?"; cat gpurun_out/r2_narrow1.json
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 900 -p no:cacheprovider -k "full_size" > gpurun_out/r2_pytest_fullsize.log 2>&1
echo "pytest full-size exit $?"; tail -8 gpurun_out/r2_pytest_fullsize.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench exit $?"; cut -c1-250 gpurun_out/r2_bench3.json; tail -3 gpurun_out/r2_bench3.err
