#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/multi_model.py 8 > gpurun_out/r2_multi_model.log 2>&1; echo "exit $?"; cat gpurun_out/r2_multi_model.log | tail -20
