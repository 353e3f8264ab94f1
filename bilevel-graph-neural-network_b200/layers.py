"""NodeEmbedding / Loss with the reference's constructor arguments, call contract and
state_dict layout (model/layers.py:9-89), computed by the sm_100a kernels."""
import torch
import torch.nn as nn

from . import ops, dist as bdist
from .config import get_flags
from .layers_util import create_act, apply_linear_act

# softmax grouping index of GATConv: 'source' = torch-geometric 1.1.x (the reference's pin,
# Dockerfile:32), 'target' = >= 1.2.  Unverifiable offline -> switchable.
GAT_SOFTMAX_GROUP = 'source'



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


class GATConv(nn.Module):
    """PyG 1.1.2 GATConv parameters, heads=1: weight [in,out], att [1,1,2*out] glorot, bias zeros."""

    def __init__(self, in_channels, out_channels, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.att = nn.Parameter(torch.empty(1, 1, 2 * out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        _glorot(self.weight)
        _glorot(self.att)


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
            self.conv = GATConv(in_dim, out_dim)
            self.act = create_act(act, out_dim)
        else:
            raise ValueError('Unknown node embedding layer type {}'.format(type))
        self.bn = bn
        if self.bn:
            self.bn = nn.BatchNorm1d(out_dim)

    def forward(self, ins, batch_data, model):
        if self.higher_level:
            graph = batch_data.merge_higher_level['merge']
            seg, S = graph.bn_row_ptr, 1
        else:
            graph = batch_data.merge_data['merge']
            seg, S = graph.chunk_row_ptr, graph.S
        csr = graph.csr
        a = self.act.code                   # None for PReLU (not fusable: it has a parameter)
        parted = getattr(graph, 'partitioned', False)     # this rank's rows of a row-partitioned upper level
        if parted and self.type != 'gcn':
            raise NotImplementedError('row-partitioned upper level: only GCN layers are partitioned; run the '
                                      '{} upper level replicated (BiGNNEngine(partition_upper=False))'.format(self.type))
        # the activation's derivative is folded into the backward of the BatchNorm that follows it (one pass less
        # over [rows, C]); only when the activation output feeds nothing but that BatchNorm
        fuse = bool(self.bn) and self.training and a is not None and a != 0 and self.type in ('gcn', 'gin')
        in_act = a if fuse else 0
        if parted:
            h = ops.linear_act(ins, self.conv.weight, None, 0, 'io')           # own rows only
            x = bdist.gcn_propagate_rows(h, self.conv.bias, graph, a if a is not None else 0, fuse)
            if a is None:
                x = self.act(x)
        elif self.type == 'gcn':
            h = ops.linear_act(ins, self.conv.weight, None, 0, 'io')
            x = ops.gcn_propagate(h, self.conv.bias, csr, a if a is not None else 0, fuse)
            if a is None:
                x = self.act(x)
        elif self.type == 'gat':
            h = ops.linear_act(ins, self.conv.weight, None, 0, 'io')
            x = ops.gat_conv(h, self.conv.att, self.conv.bias, csr, self.conv.negative_slope, GAT_SOFTMAX_GROUP)
            x = ops.activation(x, a) if a is not None else self.act(x)
        else:
            z = ops.gin_aggregate(ins, csr, self._eps_value())
            lin1, lin2 = self.conv.nn[0], self.conv.nn[2]
            # Linear -> act -> Linear (model/layers.py:27-29): the inner activation's derivative is applied in the
            # epilogue of the second Linear's backward-input GEMM (t is its saved input)
            inner = a if (a is not None and a != 0) else 0
            t = apply_linear_act(z, lin1, self.act, act_bwd_by_consumer=bool(inner))
            x = apply_linear_act(t, lin2, self.act, act_bwd_by_consumer=fuse, input_act=inner)
        if self.bn and parted and self.training:
            x = bdist.rows_batch_norm(x, self.bn.weight, self.bn.bias, graph, self.bn.running_mean,
                                      self.bn.running_var, self.bn.num_batches_tracked, self.bn.eps,
                                      self.bn.momentum, in_act)
        elif self.bn:
            sink = getattr(graph, 'bn_stats_sink', None)
            if self.training and sink is not None:
                # chunks sharded over several GPUs: keep the per-chunk statistics; the engine replays
                # the running-buffer updates in global chunk order after exchanging them
                stats = torch.empty((2, S, x.shape[1]), dtype=torch.float64, device=x.device)
                x = ops.seg_batch_norm(x, self.bn.weight, self.bn.bias, seg, S, None, None, None,
                                       self.bn.eps, self.bn.momentum, stats, in_act)
                sink.append((self.bn, stats))
            elif self.training:
                x = ops.seg_batch_norm(x, self.bn.weight, self.bn.bias, seg, S, self.bn.running_mean,
                                       self.bn.running_var, self.bn.num_batches_tracked,
                                       self.bn.eps, self.bn.momentum, None, in_act)
            else:
                x = ops.bn_eval(x, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var,
                                self.bn.eps)
        model.store_layer_output(self, x)
        if self.normalize:
            x = ops.row_normalize(x)
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

    def forward(self, ins, batch_data, _):
        fused = getattr(batch_data, 'fused_loss', None)
        if fused is not None and fused[0] is ins:          # computed by the fused pair decoder on these very scores
            batch_data.fused_loss = None
            return fused[1]
        if self.type == 'CE':
            return ops.cross_entropy(ins, batch_data.y_true_device(as_int=True))
        y_true = batch_data.y_true_device()
        return ops.bce(ins.view(-1), y_true, logits=(self.type == 'BCEWithLogits'))
