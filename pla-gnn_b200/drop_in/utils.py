'''
Utils — drop-in for PLA-GNN's code/utils.py: create_graph with the reference's signature, a `dgl` name
(main_normal.py calls dgl.seed through `from utils import *`) and the unused helpers' names.
'''
import json

import numpy as np
import torch as th

import plagnn_b200
from plagnn_b200 import dgl_shim as dgl  # noqa: F401
from plagnn_b200.utils import create_graph  # noqa: F401


def data_normalize(mat):
    """Column-wise z-score (never called by the reference's hot path; kept for API completeness)."""
    return (mat - mat.mean(0)) / mat.std(0)
