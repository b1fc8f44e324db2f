"""The reference's UNCHANGED driver (main_normal.py -> train.train, code/train.py:141-358) on the CUDA drop-ins.

Needs the reference's own scripts: /root/reference/code in the build container, or baseline/_ref/code (git-ignored; staged by
__graft_entry__.build() when /root/reference exists, so that it travels to the GPU box with the snapshot).  Skipped when neither
is there.  Nothing of the reference is edited: tools/run_reference.py puts the drop-in `model`, `utils`, `dgl` modules in
sys.modules before the script starts."""
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_code_dir():
    for cand in ("/root/reference/code", os.path.join(ROOT, "baseline", "_ref", "code")):
        if os.path.exists(os.path.join(cand, "train.py")) and os.path.exists(os.path.join(cand, "main_normal.py")):
            return cand
    return None


def test_unchanged_driver_runs_on_the_drop_ins(cuda, tmp_path):
    ref = _reference_code_dir()
    if ref is None:
        pytest.skip("the reference's scripts are not available on this machine")
    sys.path.insert(0, ROOT)
    from plagnn_b200 import synth
    n = 1200
    synth.write_reference_tree(str(tmp_path), "GSE74572", n, 30000, seed=70)
    code = tmp_path / "code"
    for name in ("train.py", "main_normal.py", "main_inter.py"):
        shutil.copy(os.path.join(ref, name), code / name)            # byte-for-byte, into the scratch tree only
    for state, script in (("normal", "main_normal.py"), ("perturbation", "main_inter.py")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference.py"), "--ref", str(code), script,
                            "-data", "GSE74572", "-f", "2", "-e", "4", "-d", "cuda"], cwd=str(code), capture_output=True, text=True,
                           timeout=1500)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        logs = sorted(glob.glob(str(tmp_path / "data" / "log" / "GSE74572" / state / "*_loc_logits.npy")))
        assert len(logs) == 20                                        # 10 fold seeds x 2 folds (train.py:162,289)
        for path in logs[:3]:
            m = np.load(path)
            assert m.shape == (n, 12) and m.dtype == np.float32 and np.isfinite(m).all() and (m > 0).all() and (m < 1).all()
    assert os.path.exists(tmp_path / "data" / "log" / "GSE74572" / "normal" / "log.tsv")
