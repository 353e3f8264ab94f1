"""NodeEmbedding / Loss with the reference's constructor arguments, call contract and
state_dict layout (model/layers.py:9-89), computed by the sm_100a kernels."""
import torch
import torch.nn as nn

from . import ops
from .config import get_flags
from .layers_util import create_act


class GINConv(nn.Module):
    """Parameter container laid out like PyG 1.1.2 GINConv: `nn` = Sequential(Linear, act,
    Linear), `eps` buffer (state_dict keys conv.nn.0.*, conv.nn.2.*, conv.eps)."""

    def __init__(self, mlps, eps=0.0):
        super().__init__()
        self.nn = mlps
        self.register_buffer('eps', torch.Tensor([eps]))


def _glorot(t):
    bound = (6.0 / (t.size(-2) + t.size(-1))) ** 0.5
    with torch.no_grad():
        t.uniform_(-bound, bound)


class GCNConv(nn.Module):
    """PyG 1.1.2 GCNConv parameters: weight [in,out] glorot, bias zeros."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        _glorot(self.weight)


class NodeEmbedding(nn.Module):
    def __init__(self, type, in_dim, out_dim, act, bn, normalize, higher_level=False, use_edge_attr=None):
        super().__init__()
        self.normalize = normalize
        self.type = type
        self.out_dim = out_dim
        self.higher_level = higher_level
        self.use_edge_attr = use_edge_attr
        if type == 'gcn':
            self.conv = GCNConv(in_dim, out_dim)
            self.act = create_act(act, out_dim)
        elif type == 'gin':
            self.act = create_act(act, out_dim)
            self.conv = GINConv(nn.Sequential(nn.Linear(in_dim, out_dim), self.act, nn.Linear(out_dim, out_dim)))
        elif type == 'gat':
            raise NotImplementedError('gat is built in a later step of the path')
        else:
            raise ValueError('Unknown node embedding layer type {}'.format(type))
        self.bn = bn
        if self.bn:
            self.bn = nn.BatchNorm1d(out_dim)
        if normalize:
            raise NotImplementedError('normalize=True is not on the Bi-GNN default path')

    def forward(self, ins, batch_data, model):
        if self.higher_level:
            graph = batch_data.merge_higher_level['merge']
            seg, S = graph.bn_row_ptr, 1
        else:
            graph = batch_data.merge_data['merge']
            seg, S = graph.chunk_row_ptr, graph.S
        csr = graph.csr
        a = self.act.code
        if self.type == 'gcn':
            h = ops.linear_act(ins, self.conv.weight, None, 0, 'io')
            x = ops.gcn_propagate(h, self.conv.bias, csr, a)
        else:
            z = ops.gin_aggregate(ins, csr, self._eps_value())
            lin1, lin2 = self.conv.nn[0], self.conv.nn[2]
            t = ops.linear_act(z, lin1.weight, lin1.bias, a, 'oi')
            x = ops.linear_act(t, lin2.weight, lin2.bias, a, 'oi')
        if self.bn:
            sink = getattr(graph, 'bn_stats_sink', None)
            if self.training and sink is not None:
                # chunks sharded over several GPUs: keep the per-chunk statistics; the engine replays
                # the running-buffer updates in global chunk order after exchanging them
                stats = torch.empty((2, S, x.shape[1]), dtype=torch.float64, device=x.device)
                x = ops.seg_batch_norm(x, self.bn.weight, self.bn.bias, seg, S, None, None, None,
                                       self.bn.eps, self.bn.momentum, stats)
                sink.append((self.bn, stats))
            elif self.training:
                x = ops.seg_batch_norm(x, self.bn.weight, self.bn.bias, seg, S, self.bn.running_mean,
                                       self.bn.running_var, self.bn.num_batches_tracked,
                                       self.bn.eps, self.bn.momentum)
            else:
                x = ops.bn_eval(x, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var,
                                self.bn.eps)
        model.store_layer_output(self, x)
        return x

    def _eps_value(self):
        # eps is a constant buffer (train_eps=False): read it once, not per step (a .item()
        # is a device sync); a state_dict load re-reads it
        if getattr(self, '_eps_cached', None) is None:
            self._eps_cached = float(self.conv.eps.item())
        return self._eps_cached

    def _load_from_state_dict(self, *args, **kwargs):
        self._eps_cached = None
        return super()._load_from_state_dict(*args, **kwargs)


class Loss(nn.Module):
    def __init__(self, type):
        super().__init__()
        self.type = type
        if type not in ('BCE', 'BCEWithLogits', 'CE'):
            raise ValueError('Unknown loss layer type {}'.format(type))
        if type == 'CE':
            raise NotImplementedError('CE (DrugCombo multi-class) is built in a later step of the path')

    def forward(self, ins, batch_data, _):
        y_true = batch_data.y_true_device()
        return ops.bce(ins.view(-1), y_true, logits=(self.type == 'BCEWithLogits'))
