"""Hands the upper level its graph (the reference's `LoadInteractionLayer`,
model/layers_load_interaction_graph.py:7-21).

The reference rebuilds a PyG `Data` object from networkx on every forward (sorting and coalescing
57 k edges per step although the train graph never changes).  Here the interaction CSR -- and, for
DrugCombo, one CSR per interaction edge type -- is resident in HBM; the layer only publishes it on
the batch and returns the current node features (the pooled drug embeddings `init_x`)."""
import torch.nn as nn

from .config import get_flags


class LoadInteractionGraph(nn.Module):
    def forward(self, x, batch_data, model):
        graph = getattr(batch_data, 'interaction_combo_nxgraph', None)
        if graph is None:
            raise AssertionError('the batch carries no interaction graph')
        upstairs = batch_data.merge_higher_level
        upstairs['merge'] = graph
        if get_flags().different_edge_type_aggr:
            upstairs['edges'] = batch_data.dataset.interaction_nxgraphs       # {edge type: graph}
        return graph.init_x
