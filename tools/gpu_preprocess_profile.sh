#!/bin/bash
# Preprocessing kernels: GPU tests, event timings at the PPI shape, then ncu --set full of one launch of
# each kernel at a size whose device memory ncu can save/restore quickly (8 192 nodes: 2 x 0.5 GB dense matrices).
# Round 1 ran the first two steps and ran out of GPU budget inside the third (only ecc_kernel was captured).
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_preprocess.py -q 2>&1 | tail -5
timeout 60 python tools/preprocess_time.py > gpurun_out/pp_time.json 2> gpurun_out/pp_time.err; cat gpurun_out/pp_time.json
timeout 60 python tools/preprocess_time.py --nodes 8192 --edges 400000 --once > gpurun_out/pp_once.json 2> gpurun_out/pp_once.err || exit 1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"ecc_kernel|pp_moment_kernel|pp_rewire_kernel|pp_emit_kernel|pp_bitmask_kernel" -c 10 \
    -o gpurun_out/prof_preprocess -f python tools/preprocess_time.py --nodes 8192 --edges 400000 --once > gpurun_out/pp_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/pp_ncu.log
