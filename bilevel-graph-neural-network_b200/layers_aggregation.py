"""Readout layers (model/layers_aggregation.py:10-75) over the segment-readout kernel."""
import torch
import torch.nn as nn

from . import ops
from .config import get_flags
from .layers_util import MLP


class NodeAggregationPairs(nn.Module):
    def __init__(self, style, concat_multi_scale=False, in_dim=None, out_dim=None, num_mlp_layers=None):
        super().__init__()
        self.style = style
        self.concat_multi_scale = concat_multi_scale
        if style in ('avg_pool', 'sum'):
            self.agg_func = None
        elif style == 'deepsets':
            self.agg_func = DeepSets(in_dim, out_dim, num_mlp_layers)
        elif style == 'gmn_aggr':
            self.agg_func = GMNAggregatorPairs(in_dim, out_dim, bn=True)
        else:
            raise NotImplementedError('{} is not implemented'.format(style))

    def forward(self, ins, batch_data, model, pair_batch=False, dst_row=None, out_rows=None):
        g = batch_data.merge_data['merge']
        acts = list(model.acts[1:]) if self.concat_multi_scale else [ins]
        if self.agg_func is None:
            return ops.readout(acts, g.seg_ptr, g.G, self.style, dst_row, out_rows)
        outs = [self.agg_func(a, g) for a in acts]
        out = torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]
        if dst_row is not None:
            full = torch.zeros((out_rows, out.shape[1]), dtype=out.dtype, device=out.device)
            out = full.index_copy(0, dst_row.long(), out)
        return out


class DeepSets(nn.Module):
    """rho(mean_pool(phi(x)))  (model/layers_aggregation.py:45-56)."""

    def __init__(self, in_dim, out_dim, num_mlp_layers):
        super().__init__()
        self.phi = MLP(in_dim, out_dim, num_hidden_lyr=num_mlp_layers - 1)
        self.rho = MLP(out_dim, out_dim, num_hidden_lyr=num_mlp_layers - 1)

    def forward(self, ins, graph):
        h = self.phi(ins)
        h = ops.readout([h], graph.seg_ptr, graph.G, 'avg_pool')
        return self.rho(h)


class GMNAggregatorPairs(nn.Module):
    """Gated ("attention") readout: MLP(sum_atoms sigmoid(gate(x)) * weight(x))
    (model/layers_aggregation.py:78-95); its MLPs carry BatchNorm, whose batches are the
    chunk's atoms (weight/gate) and the chunk's graphs (mlp_graph)."""

    def __init__(self, input_dim, output_dim, bn):
        super().__init__()
        self.out_dim = output_dim
        self.weight_func = MLP(input_dim, output_dim, num_hidden_lyr=1, hidden_channels=[output_dim], bn=bn)
        self.gate_func = MLP(input_dim, output_dim, num_hidden_lyr=1, hidden_channels=[output_dim], bn=bn)
        self.mlp_graph = MLP(output_dim, output_dim, num_hidden_lyr=1, hidden_channels=[output_dim], bn=bn)

    def forward(self, x, graph):
        atoms = (graph.chunk_row_ptr, graph.S)
        w = self.weight_func(x, seg=atoms)
        gate = self.gate_func(x, seg=atoms)
        emb = ops.gated_readout(gate, w, graph.seg_ptr, graph.G)       # sigmoid(gate) * w summed per graph, one launch
        return self.mlp_graph(emb, seg=(graph.chunk_graph_ptr, graph.S))


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
