"""MetaLayerWrapper with the per-edge-type aggregation node model (model/layers_meta.py:11-47,
61-79): the DrugCombo upper level -- one GCN/GAT per interaction edge type (synergy, antagonism),
summed, then the activation; no BatchNorm.  The edge-attribute variants (mlp_concat / gcn_concat /
gat_concat, off by default: src/config.py:77) are outside the path."""
import torch
import torch.nn as nn

from . import ops
from . import layers as L
from .layers_util import create_act


class MetaLayer(nn.Module):
    """torch_geometric.nn.MetaLayer as a parameter container (state_dict keys
    `meta_layer.node_model.*` alias `node_model.*`, as in the reference)."""

    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model, self.node_model, self.global_model = edge_model, node_model, global_model


class NodeModelAggrByEdge(nn.Module):
    def __init__(self, type, input_dim, num_edge_types, output_dim):
        super().__init__()
        self.type, self.input_dim, self.output_dim = type, input_dim, output_dim
        if type == 'gcn':
            self.GNNS = nn.ModuleList([L.GCNConv(input_dim, output_dim) for _ in range(num_edge_types - 1)])
        elif type == 'gat':
            self.GNNS = nn.ModuleList([L.GATConv(input_dim, output_dim) for _ in range(num_edge_types - 1)])
        else:
            raise ValueError

    def forward(self, x, batch_data):
        stack = getattr(batch_data.dataset, 'interaction_stack', None)
        T = len(self.GNNS)
        if self.type == 'gat' and stack is not None and T > 1 and stack.n == T * x.shape[0]:
            # batched over edge types: one stacked transform buffer, one block-diagonal GAT launch
            N = x.shape[0]
            h = ops.stacked_linear(x, [c.weight for c in self.GNNS])
            att = torch.cat([c.att.reshape(1, -1) for c in self.GNNS], 0)
            bias = torch.stack([c.bias for c in self.GNNS], 0)
            o = ops.gat_conv(h, att, bias, stack.csr, self.GNNS[0].negative_slope, L.GAT_SOFTMAX_GROUP, n_block=N)
            outs = o[:N]
            for t in range(1, T):
                outs = ops.add(outs, o[t * N:(t + 1) * N])
            return outs
        outs = None
        for i, g in enumerate(batch_data.merge_higher_level['edges'].values()):
            conv = self.GNNS[i]
            h = ops.linear_act(x, conv.weight, None, 0, 'io')
            if self.type == 'gcn':
                o = ops.gcn_propagate(h, conv.bias, g.csr, 0)
            else:
                o = ops.gat_conv(h, conv.att, conv.bias, g.csr, conv.negative_slope, L.GAT_SOFTMAX_GROUP)
            outs = o if outs is None else ops.add(outs, o)
        return outs


class MetaLayerWrapper(nn.Module):
    def __init__(self, input_dim, edge_dim, output_dim, edge_model, node_model, act, num_edge_types,
                 higher_level=True):
        super().__init__()
        self.higher_level = higher_level
        self.activation = create_act(act)
        self.num_edge_types = num_edge_types
        self.out_dim = output_dim
        if edge_model != 'none':
            raise NotImplementedError('edge-attribute MetaLayer variants are off by default and outside the path')
        self.edge_model = None
        if 'multi_edge_aggr' not in node_model:
            raise NotImplementedError('only the *_multi_edge_aggr node models are on the path')
        assert higher_level
        self.node_model = NodeModelAggrByEdge(node_model.split('_')[0], input_dim, num_edge_types, output_dim)
        self.meta_layer = MetaLayer(self.edge_model, self.node_model)

    def forward(self, ins, batch_data, model):
        out = self.node_model(ins, batch_data)
        code = self.activation.code
        return ops.activation(out, code) if code is not None else self.activation(out)
