"""BASELINE.json configs[1] end to end on the device: train the control state and the perturbation state, merge the
per-model output matrices of each state, score the alteration against the control and rank it.

What it mirrors in the reference (paths relative to the reference checkout):

* ``code/train.py:162-205,289``  the sweep of one state: for each of the fold seeds a ``KFold(fold_num, shuffle=True,
  random_state=seed)`` over the labelled nodes, a fresh ``GNN32(F, 400, 300, 200, 100, 12)`` + ``Adam(lr)`` per fold,
  ``epoch_num`` epochs, and the LAST epoch's pre-step output saved per (round, fold) (``np.save(..._loc_logits)``);
* ``code/main.py:32-48``  ``mat_merge``: mean over the saved matrices of ``scaling(mat)``;
* ``code/main.py:80-84,143-175``  ``(inter - normal) / normal`` on the scaled matrices and the descending rank.

Everything between the inputs and the ranked table runs in the library's kernels: each model trains through ``TrainStep``
(one CUDA-graph replay per epoch), the output matrices never leave the device, merge and ranking are ``scoring.py``.
The reference's logging / per-epoch metrics (``train.py:206-279``) are not part of this path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import scoring
from .epoch import TrainStep
from .loss import weight_cal
from .nn import GNN32

FOLD_SEEDS = (12, 22, 32, 42, 52, 62, 72, 82, 92, 100)      # code/train.py:162
HIDDEN = (400, 300, 200, 100, 12)                           # code/train.py:179


def fold_splits(label_index, fold_num: int, seed: int):
    """The (train_index, val_index) node-id lists of code/train.py:165,178-188 for one fold seed."""
    from sklearn.model_selection import KFold
    label = np.asarray(label_index)
    for train_idx, val_idx in KFold(n_splits=fold_num, random_state=seed, shuffle=True).split(label):
        yield label[train_idx], label[val_idx]


def train_state(g, label_index, loc_mat, lr: float = 5e-5, fold_num: int = 10, epoch_num: int = 200, fold_seeds=FOLD_SEEDS,
                model_seed: int | None = None, use_graph: bool = True, on_model=None, max_models: int | None = None):
    """One state's sweep (code/train.py:141-289 without its logging): returns the list of last-epoch output matrices
    (device tensors, N x 12), one per (fold seed, fold), in the reference's order."""
    features, labels = g.ndata["feat"], g.ndata["loc"]
    i_weight = weight_cal(np.asarray(loc_mat))
    if model_seed is not None:
        torch.manual_seed(model_seed)
    outs = []
    for seed in fold_seeds:
        for train_index, _val_index in fold_splits(label_index, fold_num, seed):
            model = GNN32(features.shape[1], *HIDDEN).to(features.device)
            step = TrainStep(model, g, features, labels, train_index, i_weight, lr=lr, use_graph=use_graph)
            step.run(epoch_num)
            outs.append(step.logits.clone())            # the last epoch's pre-step output (train.py:289 saves `logits`)
            if on_model is not None:
                on_model(model, step)
            del step, model
            if max_models is not None and len(outs) >= max_models:      # a bounded sample of the sweep (bench.py)
                return outs
    return outs


def alteration_pipeline(g_normal, g_inter, label_index, loc_mat, lr: float = 5e-5, fold_num: int = 10, epoch_num: int = 200,
                        fold_seeds=FOLD_SEEDS, model_seed: int | None = 70, top: int | None = None, use_graph: bool = True,
                        max_models: int | None = None):
    """main_normal.py + main_inter.py + main.py's scoring as one call.  Returns (records, normal_merged, inter_merged):
    records = scoring.misloc_records(...) (row, col, score, normal, perturbation, rank as device tensors)."""
    normal_outs = train_state(g_normal, label_index, loc_mat, lr, fold_num, epoch_num, fold_seeds, model_seed, use_graph,
                              max_models=max_models)
    inter_outs = train_state(g_inter, label_index, loc_mat, lr, fold_num, epoch_num, fold_seeds, model_seed, use_graph,
                             max_models=max_models)
    normal_merged = scoring.mat_merge(normal_outs)
    inter_merged = scoring.mat_merge(inter_outs)
    return scoring.misloc_records(normal_merged, inter_merged, top=top), normal_merged, inter_merged
