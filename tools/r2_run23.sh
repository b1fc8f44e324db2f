#!/bin/bash
mkdir -p gpurun_out
export PLAGNN_LIB_PATH=pla-gnn_b200/libplagnn_diag.so
export PLAGNN_TMA_DB_NOW=1
for dbg in 0 5 6 2 7; do echo "== debug $dbg"; PLAGNN_TMA_DEBUG=$dbg timeout 120 python tools/gemm_trace_db.py 24041 503 503 0 2>&1 | tail -3; done > gpurun_out/r2_gemm_trace_db_modes.log 2>&1
cat gpurun_out/r2_gemm_trace_db_modes.log
