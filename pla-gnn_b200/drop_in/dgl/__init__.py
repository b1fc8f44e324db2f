"""Minimal `dgl` stand-in exposing only what PLA-GNN touches (utils.py:44-49, main_normal.py:15, model.py:7)."""
from plagnn_b200.dgl_shim import *  # noqa: F401,F403
from plagnn_b200.dgl_shim import graph, add_self_loop, seed  # noqa: F401
from . import nn  # noqa: F401
