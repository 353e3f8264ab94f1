"""LoadInteractionGraph (model/layers_load_interaction_graph.py:7-21): hands the upper
level its graph.  The reference rebuilds a PyG `Data` from networkx every forward; the
train interaction graph never changes, so the CSR lives in HBM and only `x` is attached."""
import torch.nn as nn

from .config import get_flags


class LoadInteractionGraph(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, x, batch_data, model):
        assert hasattr(batch_data, 'interaction_combo_nxgraph')
        ig = batch_data.interaction_combo_nxgraph
        batch_data.merge_higher_level['merge'] = ig
        if get_flags().different_edge_type_aggr:
            # one graph per interaction edge type (layers_load_interaction_graph.py:16-19); static CSRs here
            batch_data.merge_higher_level['edges'] = batch_data.dataset.interaction_nxgraphs
        return ig.init_x
