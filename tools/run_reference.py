#!/usr/bin/env python
"""Runs the UNCHANGED reference scripts (main_normal.py / main_inter.py -> train.py) on the CUDA drop-ins.

    python tools/run_reference.py --ref /path/to/PLA-GNN/code main_normal.py -data GSE74572 -f 2 -e 3

The working directory must be the reference's code/ directory of a tree that has data/generate_materials/
(plagnn_b200.synth.write_reference_tree writes a synthetic one).  Nothing in the reference is edited: the
drop-in `model`, `utils`, `dgl`, `matplotlib` modules are placed in sys.modules before the script starts, and
train.protein_loc_correction (a 24 041-iteration Python loop that also indexes a CPU tensor with a CUDA mask) is
replaced by the on-device kernel unless --keep-python-metrics is given.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True, help="path to the reference's code/ directory")
    ap.add_argument("--keep-python-metrics", action="store_true")
    ap.add_argument("--fused-loss-adam", action="store_true", help="also swap train.multi_loss / torch.optim.Adam for the fused kernels")
    ap.add_argument("script")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    import statistics  # noqa: F401  (stdlib; the reference ships a same-named script)
    import torch._dynamo  # noqa: F401
    drop = os.path.join(ROOT, "pla-gnn_b200", "drop_in")
    sys.path.insert(0, drop)
    import model, utils, dgl, matplotlib.pyplot  # noqa: F401,E401  -> the drop-ins are now cached in sys.modules
    sys.path.append(a.ref)
    os.chdir(a.ref)
    import train  # the reference's train.py, unmodified
    import plagnn_b200 as P
    if not a.keep_python_metrics:
        train.protein_loc_correction = lambda proba, alpha: P.protein_loc_correction(proba, alpha)
        train.performances_record = P.performances_record
    if a.fused_loss_adam:
        import torch
        train.multi_loss = P.multi_loss
        torch.optim.Adam = P.FusedAdam
    sys.argv = [os.path.join(a.ref, a.script)] + a.rest
    runpy.run_path(os.path.join(a.ref, a.script), run_name="__main__")


if __name__ == "__main__":
    main()
