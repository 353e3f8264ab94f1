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

    def _fused_head(self, ins, model, batch_data):
        """(head code, target tensor or None) when the fused decoder kernel covers this scorer + the loss layer that
        follows it; None -> the layer-by-layer path (BIGNN_NO_FUSED_DECODER=1 forces it)."""
        import os
        if self.type != 'mlp_concat' or os.environ.get('BIGNN_NO_FUSED_DECODER') or not ins.is_cuda:
            return None
        mlp = self.mlp_concat
        if mlp.bn or mlp.activation.code != ops.ACT_CODES['relu']:
            return None
        if not ops.pair_decoder_supported(ins.shape[1], [l.out_features for l in mlp.layers]):
            return None
        if any(l.bias is None for l in mlp.layers):
            return None
        loss_layer = model.layers[-1] if getattr(model, 'layers', None) is not None else None
        kind = getattr(loss_layer, 'type', None)
        if self.multi_label_pred:
            if kind != 'CE':
                return None
            head = 2
        else:
            if kind not in ('BCE', 'BCEWithLogits'):
                return None
            # the scorer itself applies the sigmoid (layers_link_pred.py:65): the Loss layer sees probabilities for
            # 'BCE'; with 'BCEWithLogits' the reference feeds those probabilities to the logits loss -- not fusable
            if kind != 'BCE':
                return None
            head = 0
        if not model.training or not hasattr(batch_data, 'y_true_device'):
            return head, None
        try:
            target = batch_data.y_true_device(as_int=(head == 2))
        except Exception:
            return head, None
        return head, target

    def forward(self, ins, batch_data, model):
        graph = getattr(batch_data, 'merge_higher_level', {}).get('merge')
        if getattr(graph, 'partitioned', False):
            # row-partitioned upper level: all-gather the final embeddings; the pair batch is scored by every
            # rank (P pairs << nnz), pair rows are positions of the gathered matrix (engine.stage_pairs)
            ins = bdist.gather_rows_for_replicated_consumer(ins, graph)
        rows, entry_csr = batch_data.pair_rows_device(ins.shape[0], higher=get_flags().higher_level_layers,
                                                      unique=self.batch_unique_graphs)
        fused = self._fused_head(ins, model, batch_data)
        if fused is not None:
            # gather + normalise + concat + MLP + head + LOSS in one launch (forward) / one launch (backward); the
            # Loss layer that follows picks the loss up from the batch
            head, target = fused
            lins = self.mlp_concat.layers
            params = [t for lin in lins for t in (lin.weight, lin.bias)]
            scores, loss = ops.pair_decoder(ins, rows, entry_csr, head, target, params)
            batch_data.fused_loss = (scores, loss) if target is not None else None
            batch_data.assign_link_preds(scores)
            return scores
        z = ops.pair_gather_norm(ins, rows, entry_csr)                 # [P, 2D] = [norm(h_a) || norm(h_b)]
        sigmoid = ops.ACT_CODES['sigmoid']
        if self.type == 'dot_product':
            scores = ops.pair_dot(z, sigmoid)
        else:
            scores = self.mlp_concat(z, final_act=0 if self.multi_label_pred else sigmoid)
        batch_data.assign_link_preds(scores)
        return scores
