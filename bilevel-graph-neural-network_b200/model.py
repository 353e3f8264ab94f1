"""Model shell (model/model.py:7-82): sequential layer list, `acts` bookkeeping and the
`use_layers` switch, over the layers built by this package's registry."""
import torch.nn as nn

from .config import get_flags
from .layers_factory import create_layers


class Model(nn.Module):
    def __init__(self, data, config_layer_type='layer'):
        super().__init__()
        flags = get_flags()
        self.train_data = data
        self.interaction_num_node_feat = data.dataset.interaction_num_node_feat
        self.num_node_feat = data.num_node_feat
        self.num_hyper_edge_feat = data.num_hyper_edge_feat
        self.num_labels = data.dataset.num_labels
        self.layers = create_layers(self, config_layer_type, vars(flags)['{}_num'.format(config_layer_type)])
        self.pred_layer = self.layers[-2]
        self._use_layers = 'all'
        if flags.lower_level_layers and flags.higher_level_layers:
            self._use_layers = 'init_model'
            self.init_layers = self.layers[:flags.last_lower_lyr_num]
            self.lower_layers = self.layers[:flags.last_lower_lyr_num + 1]
            self.higher_level_layers = self.layers[flags.last_lower_lyr_num + 1:]
        elif flags.lower_level_layers:
            self.init_layers = self.layers[:-2]
            self.lower_layers = self.layers[:-2]
        assert len(self.layers) > 0
        self.layer_output = {}
        self.acts = None

    def _select(self):
        flags = get_flags()
        u = self._use_layers
        if flags.lower_level_layers and flags.higher_level_layers:
            if u == 'init_layers':
                return self.init_layers
            if u == 'lower_layers':
                return self.lower_layers
            if u == 'higher_layers':
                return self.higher_level_layers
            if u == 'higher_no_eval_layers':
                return self.layers[flags.last_lower_lyr_num + 1:-2]
            raise UnboundLocalError("use_layers must be set before forward (got {!r})".format(u))
        if u == 'higher_no_eval_layers':
            return self.layers[:-2]
        if u == 'lower_layers':
            return self.lower_layers
        if u == 'init_layers':
            return self.init_layers
        return self.layers

    def forward(self, batch_data):
        md = batch_data.merge_data.get('merge')
        self.acts = [md.x if md is not None else None]
        for layer in self._select():
            self.acts.append(layer(self.acts[-1], batch_data, self))
        return self.acts[-1]

    def store_layer_output(self, layer, output):
        self.layer_output[layer] = output

    def get_layer_output(self, layer):
        return self.layer_output[layer]

    @property
    def use_layers(self):
        return self._use_layers

    @use_layers.setter
    def use_layers(self, setting):
        assert setting in ['all', 'init_layers', 'lower_layers', 'higher_layers', 'higher_no_eval_layers']
        self._use_layers = setting
