'''
Models — drop-in for PLA-GNN's code/model.py (same names, same signatures), backed by plagnn_b200.
Copy this file over code/model.py (or put this directory first on sys.path).
'''
import torch as th
import torch.nn as nn
import torch.nn.functional as F

import plagnn_b200
from plagnn_b200.nn import SAGEConv, GNN32  # noqa: F401  (train.py does `from model import *`)
