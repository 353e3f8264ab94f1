"""bignn_b200 -- the B200-native Bi-GNN bi-level message-passing path.

Host-side mirror of the reference's operator API for that path (model/layers_factory.py
`layer_ctors` + the `layer(ins, batch_data, model)` call contract) over a C-ABI CUDA
library for sm_100a (include/bignn_b200.h).  Import name: `bignn_b200` (the directory
name carries the reference repository's name and is not a Python identifier).
"""
from . import _lib
from .config import make_flags, set_flags, get_flags
from .dataset import BiGNNData
from .graph import PackedGraphs, MergedGraph, InteractionGraph
from .batch import BatchData, sample_negative_pairs
from .sampler import RandomSampler, EverythingSampler, NeighborSampler
from .layers_factory import create_layers, layer_ctors
from .model import Model
from . import ops, train

__all__ = ['make_flags', 'set_flags', 'get_flags', 'BiGNNData', 'PackedGraphs', 'MergedGraph',
           'InteractionGraph', 'BatchData', 'sample_negative_pairs', 'RandomSampler', 'EverythingSampler', 'NeighborSampler', 'create_layers',
           'layer_ctors', 'Model', 'ops', 'train']
