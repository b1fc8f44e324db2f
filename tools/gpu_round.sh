#!/bin/bash
# One GPU-box session: tests, diagnostics, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== tests, tcgen05 default"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?"; tail -12 gpurun_out/pytest_gpu.log
echo "== tcgen05 diagnostic"
timeout 300 python tools/diag_tcgen05.py > gpurun_out/diag_tcgen05.log 2>&1
echo "exit $?"; grep -c "rel err" gpurun_out/diag_tcgen05.log; tail -2 gpurun_out/diag_tcgen05.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench (tcgen05)"
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("gemm", d["gemm"]); print("spmm", d["spmm"]); print("cpu", d.get("cpu_baseline"))
for k in json.load(open("gpurun_out/bench_kernels_n1.json")): print(k)
PY
tail -5 gpurun_out/bench.err
echo "== bench (simt) for comparison"
PLAGNN_GEMM=simt timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err
echo "exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_simt.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['gemm'])"
