"""plagnn_b200 — B200-native (sm_100a) implementation of PLA-GNN's message-passing hot path.

Public surface (mirrors what the reference's train loop touches):
    GNN32, SAGEConv            model.py drop-ins (nn.py)
    create_graph, Graph, graph, add_self_loop    utils.py / dgl drop-ins (utils.py, graph.py)
    multi_loss, multi_loss_indexed, weight_cal   fused loss (loss.py)
    FusedAdam                  fused optimiser (optim.py)
    TrainStep                  the epoch body (forward, loss, backward, Adam) as one CUDA graph replay (epoch.py)
    protein_loc_correction     on-device label decision (metrics.py)
    pipeline                   configs[1] end to end: train both states, merge, score, rank (pipeline.py)
    scoring                    alteration scoring after training: scaling / mat_merge / alteration_rank (scoring.py)
    preprocess                 offline stage on the device: edge_clustering_coefficients / modify_network_topology (preprocess.py)
The kernels live in libplagnn.so (csrc/, C ABI in include/plagnn.h); importing this package does not
need a GPU, calling any kernel does.
"""
from . import _lib
from ._lib import PlagnnError, LIB_PATH
from .graph import Graph, Csr, graph, add_self_loop, build_csr
from .nn import GNN32, SAGEConv, GraphConvSum, GCN
from .loss import multi_loss, multi_loss_indexed, weight_cal
from .optim import FusedAdam
from .epoch import TrainStep
from .metrics import protein_loc_correction, performances_record
from .utils import create_graph
from . import scoring
from . import pipeline
from . import preprocess

__all__ = ["GNN32", "SAGEConv", "GraphConvSum", "GCN", "Graph", "Csr", "graph", "add_self_loop", "build_csr",
           "create_graph", "multi_loss", "multi_loss_indexed", "weight_cal", "FusedAdam", "TrainStep", "protein_loc_correction",
           "performances_record", "PlagnnError", "LIB_PATH"]
