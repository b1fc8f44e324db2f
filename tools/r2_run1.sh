#!/bin/bash
# round 2, first GPU call: the model tests (gradient parity on shared decisions), smoke, then the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2_pytest_model.log 2>&1
echo "pytest model exit $?"; tail -5 gpurun_out/r2_pytest_model.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_model.py > gpurun_out/r2_pytest_rest.log 2>&1
echo "pytest rest exit $?"; tail -4 gpurun_out/r2_pytest_rest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench exit $?"; cut -c1-900 gpurun_out/r2_bench1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | cut -c1-300
