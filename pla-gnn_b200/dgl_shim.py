"""The slice of the `dgl` module surface the reference uses, served by plagnn_b200.graph / nn:
dgl.graph, dgl.add_self_loop, dgl.seed, dgl.nn.pytorch.SAGEConv."""
import types

import torch

from .graph import Graph, add_self_loop, graph  # noqa: F401
from .nn import SAGEConv

DGLGraph = Graph


def seed(val):
    """dgl.seed (main_normal.py:15): DGL's own RNG is not used on this path; seed torch's for parity of intent."""
    torch.manual_seed(val)


nn = types.ModuleType("plagnn_b200.dgl_shim.nn")
nn.pytorch = types.ModuleType("plagnn_b200.dgl_shim.nn.pytorch")
nn.pytorch.SAGEConv = SAGEConv
nn.pytorch.conv = types.ModuleType("plagnn_b200.dgl_shim.nn.pytorch.conv")
nn.pytorch.conv.SAGEConv = SAGEConv
