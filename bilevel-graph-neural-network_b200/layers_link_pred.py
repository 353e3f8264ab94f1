"""LinkPred (model/layers_link_pred.py:9-71): normalise + gather + concat + MLP + sigmoid."""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import get_flags
from .graph import entry_csr
from .layers_util import MLP


class LinkPred(nn.Module):
    def __init__(self, type, mlp_dim, num_labels, weight_dim=None, batch_unique_graphs=True,
                 multi_label_pred=False):
        super().__init__()
        self.type = type
        self.num_labels = num_labels
        self.batch_unique_graphs = batch_unique_graphs
        self.weight_dim = weight_dim
        self.multi_label_pred = multi_label_pred
        if self.type not in ['dot_product', 'mlp_concat']:
            raise NotImplementedError
        if weight_dim:
            self.weight_matrix = nn.Parameter(torch.zeros((weight_dim, weight_dim)))
            nn.init.xavier_normal_(self.weight_matrix, gain=nn.init.calculate_gain('relu'))
        if self.multi_label_pred:
            assert self.type == 'ntn' or self.type == 'mlp_concat'
        if self.type != 'mlp_concat':
            pass
        elif multi_label_pred:
            dims = self._calc_mlp_dims(mlp_dim * 2, num_labels, division=8)
            self.mlp_concat = MLP(mlp_dim * 2, num_labels, num_hidden_lyr=len(dims), hidden_channels=dims, bn=False)
        else:
            dims = self._calc_mlp_dims(mlp_dim * 2, division=8)
            self.mlp_concat = MLP(mlp_dim * 2, 1, num_hidden_lyr=len(dims), hidden_channels=dims, bn=False)

    @staticmethod
    def _calc_mlp_dims(mlp_dim, output_dim=1, division=2):
        dim = mlp_dim
        dims = []
        while dim > output_dim:
            dim = dim // division
            dims.append(dim)
        return dims[:-1]

    def forward(self, ins, batch_data, model):
        ids_dev, ecsr = batch_data.pair_rows_device(ins.shape[0],
                                                    higher=get_flags().higher_level_layers,
                                                    unique=self.batch_unique_graphs)
        z = ops.pair_gather_norm(ins, ids_dev, ecsr)
        if self.type == 'dot_product':
            pair_preds = ops.pair_dot(z, ops.ACT_CODES['sigmoid'])
        else:
            final = 0 if self.multi_label_pred else ops.ACT_CODES['sigmoid']
            pair_preds = self.mlp_concat(z, final_act=final)
        batch_data.assign_link_preds(pair_preds)
        return pair_preds
