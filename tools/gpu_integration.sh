#!/bin/bash
# One-off: the reference's UNCHANGED train.py / main_normal.py (staged, git-ignored, under baseline/_ref) on the CUDA drop-ins.
mkdir -p gpurun_out
python - <<'PY'
import sys; sys.path.insert(0, ".")
from plagnn_b200 import synth
synth.write_reference_tree("baseline/_ref", "GSE74572", 3000, 90000, seed=70)
print("tree written")
PY
cd baseline/_ref/code
timeout 900 python ../../../tools/run_reference.py --ref "$PWD" main_normal.py -data GSE74572 -f 2 -e 6 -d cuda > ../../../gpurun_out/unchanged_driver.log 2>&1
echo "exit $?"
cd ../../..
grep -E "^(tra|val) --|learning rate" gpurun_out/unchanged_driver.log | head -12; tail -3 gpurun_out/unchanged_driver.log
ls baseline/_ref/data/log/GSE74572/normal | head
