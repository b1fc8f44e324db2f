#!/bin/bash
# whole GPU suite, smoke, bench line on one GPU with the session's defaults
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2_28_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2_28_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_28_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_28_smoke.log
tools/r2_run12_final_multi.sh 1
