#!/bin/bash
# round 2 ncu evidence (1 GPU): launch list of bench.py, --set full of the hot kernels at the PPI shape, of the aggregation
# kernels on the 1 M / 100 M graph, and of the preprocessing kernels.  Every program runs once without ncu first.  The reports
# are summarised ON the box (raw-page CSV + the markdown table) and removed: gpurun copies back at most 64 MiB.
mkdir -p gpurun_out
summ() {  # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2> /dev/null
  python tools/summarize_ncu.py full gpurun_out/$1.ncu-rep > gpurun_out/$1.md 2> gpurun_out/$1.md.err
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned --no-pipeline > gpurun_out/r2_plain.log 2>&1; echo "plain exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 900 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned --no-pipeline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches exit $?"
python tools/summarize_ncu.py launches gpurun_out/launches_r2.csv > gpurun_out/launches_r2.md 2>&1
timeout 300 python tools/kernels_once.py all > gpurun_out/r2_k1.log 2>&1; echo "kernels_once exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tma|spmm_kernel|spmm_max_scatter|narrow_" -c 32 -f -o gpurun_out/prof_r2_ppi python tools/kernels_once.py all > gpurun_out/r2_ncu_ppi.log 2>&1; echo "ncu ppi exit $?"
summ prof_r2_ppi
timeout 300 python tools/scaled_once.py > gpurun_out/r2_s1.log 2>&1; echo "scaled_once exit $?"
timeout 1200 ncu --set full --clock-control none -k regex:"spmm_kernel|spmm_narrow_kernel|spmm_max_scatter|spmm_combine" -c 12 -f -o gpurun_out/prof_r2_scaled python tools/scaled_once.py > gpurun_out/r2_ncu_scaled.log 2>&1; echo "ncu scaled exit $?"
summ prof_r2_scaled
timeout 120 python tools/preprocess_time.py > gpurun_out/r2_pp_time.json 2> gpurun_out/r2_pp_time.err; echo "pp timing exit $?"; cat gpurun_out/r2_pp_time.json
timeout 60 python tools/preprocess_time.py --nodes 8192 --edges 400000 --once > gpurun_out/pp_once.json 2> gpurun_out/pp_once.err; echo "pp once exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"ecc_kernel|pp_moment_kernel|pp_rewire_kernel|pp_emit_kernel|pp_bitmask_kernel|pp_pearson_kernel|pp_center_kernel" -c 14 -f -o gpurun_out/prof_r2_preprocess python tools/preprocess_time.py --nodes 8192 --edges 400000 --once > gpurun_out/r2_ncu_pp.log 2>&1; echo "ncu preprocess exit $?"
summ prof_r2_preprocess
du -sh gpurun_out; ls -la gpurun_out | head -40
