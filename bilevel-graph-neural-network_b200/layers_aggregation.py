"""Readout layers (model/layers_aggregation.py:10-75) over the segment-readout kernel."""
import torch
import torch.nn as nn

from . import ops
from .config import get_flags


class NodeAggregationPairs(nn.Module):
    def __init__(self, style, concat_multi_scale=False, in_dim=None, out_dim=None, num_mlp_layers=None):
        super().__init__()
        self.style = style
        self.concat_multi_scale = concat_multi_scale
        if style not in ('avg_pool', 'sum'):
            if style in ('deepsets', 'gmn_aggr'):
                raise NotImplementedError('{} readout is built in a later step of the path'.format(style))
            raise NotImplementedError('{} is not implemented'.format(style))

    def forward(self, ins, batch_data, model, pair_batch=False, dst_row=None, out_rows=None):
        g = batch_data.merge_data['merge']
        acts = list(model.acts[1:]) if self.concat_multi_scale else [ins]
        return ops.readout(acts, g.seg_ptr, g.G, self.style, dst_row, out_rows)


class NodeAggregation(NodeAggregationPairs):
    """One embedding per unique graph; in Bi-GNN mode the pooled rows also become the
    interaction graph's node features (layers_aggregation.py:66-75).  The reference writes
    them one row at a time from Python; here init_x is a functional row scatter so that the
    gradient of the upper level flows back to the chunk that produced each row."""

    def __init__(self, style, is_last_layer, **kwargs):
        super().__init__(style, **kwargs)
        self.is_last_layer = is_last_layer

    def forward(self, x, batch_data, model, pair_batch=False):
        out = super().forward(x, batch_data, model, pair_batch)
        if not get_flags().higher_level_layers:
            return out
        rows = batch_data.merge_data['dataset_rows']          # gs_map[gid] per merged graph, device int64
        ig = batch_data.interaction_combo_nxgraph
        ig.init_x = ig.init_x.index_copy(0, rows, out)
        return out
