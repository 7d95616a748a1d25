"""B200-native embedding + recommendation hot path of hieucnm/GNN-RecSys.

Import as ``gnn_recsys_b200`` (the repo-root alias module) -- the directory name carries a hyphen.
"""
from .graph import (HeteroGraph, Block, Relation, DeviceEdgeGraph, heterograph, edge_graph, csr_by_dst_host,  # noqa: F401
                    NID, EID)
from .synthetic import make_graph, make_graph_device, SyntheticData, CONFIGS  # noqa: F401
from .dataloading import (MultiLayerFullNeighborSampler, MultiLayerNeighborSampler, NodeDataLoader,  # noqa: F401
                          EdgeDataLoader, negative_sampler, to_block, sample_key, hash64)
from .model import (ConvModel, ConvLayer, NodeEmbedding, HeteroGraphConv, CosinePrediction,  # noqa: F401
                    max_margin_loss)
from .train.run import get_embeddings  # noqa: F401
from .metrics import (get_recs, get_recs_tensor, create_already_bought, create_already_bought_csr,  # noqa: F401
                      create_ground_truth, recs_to_metrics, get_metrics_at_k, metrics_from_tensor)
from .recs import RecsConfig, BoughtCSR, ScoringTable, recommend_topk  # noqa: F401
from . import ops, _native, distributed, recs  # noqa: F401
