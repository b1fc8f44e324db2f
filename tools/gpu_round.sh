#!/bin/bash
# One GPU-box session: tests, diagnostics, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== main tests (SIMT GEMM forced)" 
PLAGNN_GEMM=simt timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -k "not tcgen05" -p no:cacheprovider > gpurun_out/pytest_simt.log 2>&1
echo "exit $?"; tail -5 gpurun_out/pytest_simt.log
echo "== tcgen05 diagnostic"
timeout 300 python tools/diag_tcgen05.py > gpurun_out/diag_tcgen05.log 2>&1
echo "exit $?"; tail -30 gpurun_out/diag_tcgen05.log
echo "== smoke (simt)"
PLAGNN_GEMM=simt timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_simt.log 2>&1
echo "exit $?"; tail -3 gpurun_out/smoke_simt.log
echo "== bench (simt)"
PLAGNN_GEMM=simt timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 8 > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
echo "exit $?"; tail -c 3000 gpurun_out/bench_simt.json; tail -5 gpurun_out/bench_simt.err
