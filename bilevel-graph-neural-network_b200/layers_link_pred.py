"""Pair scorer with the reference's constructor contract and parameter layout
(model/layers_link_pred.py:9-71): L2-normalised drug embeddings of a pair are gathered and
concatenated by one kernel, scored by the `mlp_concat` MLP (sigmoid folded into the last GEMM) or by
a dot product; multi-class mode returns the logits."""
import torch
import torch.nn as nn

from . import ops, dist as bdist
from .config import get_flags
from .layers_util import MLP

_SCORERS = ('dot_product', 'mlp_concat')


def hidden_widths(width, out_width=1, division=2):
    """Widths of the scorer's hidden layers: keep dividing by `division` while above `out_width`,
    then drop the narrowest (128 -> [16, 2] for division 8; 640 -> [80, 10]; 128 -> [16] for 3 classes)."""
    widths = []
    while width > out_width:
        width //= division
        widths.append(width)
    return widths[:-1]


class LinkPred(nn.Module):
    def __init__(self, type, mlp_dim, num_labels, weight_dim=None, batch_unique_graphs=True,
                 multi_label_pred=False):
        super().__init__()
        if type not in _SCORERS:
            raise NotImplementedError
        if multi_label_pred and type != 'mlp_concat':
            raise AssertionError('multi-class prediction needs the mlp_concat scorer')
        self.type, self.num_labels = type, num_labels
        self.batch_unique_graphs, self.multi_label_pred = batch_unique_graphs, multi_label_pred
        self.weight_dim = weight_dim
        if weight_dim:                                  # parameter kept for checkpoint compatibility
            self.weight_matrix = nn.Parameter(torch.zeros((weight_dim, weight_dim)))
            nn.init.xavier_normal_(self.weight_matrix, gain=nn.init.calculate_gain('relu'))
        if type == 'mlp_concat':
            n_out = num_labels if multi_label_pred else 1
            hidden = hidden_widths(2 * mlp_dim, n_out, division=8)
            self.mlp_concat = MLP(2 * mlp_dim, n_out, num_hidden_lyr=len(hidden), hidden_channels=hidden, bn=False)

    _calc_mlp_dims = staticmethod(hidden_widths)

    def forward(self, ins, batch_data, model):
        graph = getattr(batch_data, 'merge_higher_level', {}).get('merge')
        if getattr(graph, 'partitioned', False):
            # row-partitioned upper level: all-gather the final embeddings; the pair batch is scored by every
            # rank (P pairs << nnz), pair rows are positions of the gathered matrix (engine.stage_pairs)
            ins = bdist.gather_rows_for_replicated_consumer(ins, graph)
        rows, entry_csr = batch_data.pair_rows_device(ins.shape[0], higher=get_flags().higher_level_layers,
                                                      unique=self.batch_unique_graphs)
        z = ops.pair_gather_norm(ins, rows, entry_csr)                 # [P, 2D] = [norm(h_a) || norm(h_b)]
        sigmoid = ops.ACT_CODES['sigmoid']
        if self.type == 'dot_product':
            scores = ops.pair_dot(z, sigmoid)
        else:
            scores = self.mlp_concat(z, final_act=0 if self.multi_label_pred else sigmoid)
        batch_data.assign_link_preds(scores)
        return scores
