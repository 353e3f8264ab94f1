"""Model shell over this package's layer registry.

Keeps the public surface of the reference's `Model` (model/model.py:7-82) that the hot loop and the
layers rely on -- `layers`, `pred_layer`, the `init_layers` / `lower_layers` / `higher_level_layers`
slices, the `use_layers` switch, `acts`, `store_layer_output` -- but is organised around a table of
named layer slices instead of the reference's if/elif chain.
"""
import torch.nn as nn

from .config import get_flags
from .layers_factory import create_layers

_VALID_SWITCH = ('all', 'init_layers', 'lower_layers', 'higher_layers', 'higher_no_eval_layers')


class Model(nn.Module):
    def __init__(self, data, config_layer_type='layer'):
        super().__init__()
        flags = get_flags()
        ds = data.dataset
        # sizes the layer constructors read from the model (layers_factory.get_input_dim_higher_level)
        self.train_data = data
        self.num_node_feat = data.num_node_feat
        self.num_hyper_edge_feat = data.num_hyper_edge_feat
        self.interaction_num_node_feat = ds.interaction_num_node_feat
        self.num_labels = ds.num_labels

        n_layers = vars(flags)[config_layer_type + '_num']
        self.layers = create_layers(self, config_layer_type, n_layers)
        if len(self.layers) == 0:
            raise AssertionError('a model needs at least one layer')
        self.pred_layer = self.layers[-2]            # ... -> LinkPredictor -> Loss

        self._bilevel = bool(flags.lower_level_layers and flags.higher_level_layers)
        self._use_layers = 'all'
        if self._bilevel:
            cut = flags.last_lower_lyr_num            # index of the readout layer
            self.init_layers = self.layers[:cut]
            self.lower_layers = self.layers[:cut + 1]
            self.higher_level_layers = self.layers[cut + 1:]
            self._use_layers = 'init_model'           # callers must pick a slice first (src/train.py:51,97)
        elif flags.lower_level_layers:
            self.init_layers = self.lower_layers = self.layers[:-2]
        self.layer_output = {}
        self.acts = None

    # ------------------------------------------------------------------ layer slices
    def _active_layers(self):
        key = self._use_layers
        if self._bilevel:
            table = {'init_layers': lambda: self.init_layers,
                     'lower_layers': lambda: self.lower_layers,
                     'higher_layers': lambda: self.higher_level_layers,
                     'higher_no_eval_layers': lambda: self.higher_level_layers[:-2]}
            if key not in table:
                raise UnboundLocalError('use_layers must be set before forward (got {!r})'.format(key))
            return table[key]()
        if key == 'higher_no_eval_layers':
            return self.layers[:-2]
        if key in ('lower_layers', 'init_layers'):
            return getattr(self, key)
        return self.layers

    @property
    def use_layers(self):
        return self._use_layers

    @use_layers.setter
    def use_layers(self, setting):
        if setting not in _VALID_SWITCH:
            raise AssertionError('use_layers must be one of {}'.format(_VALID_SWITCH))
        self._use_layers = setting

    # ------------------------------------------------------------------ forward
    def forward(self, batch_data):
        merged = batch_data.merge_data.get('merge')
        acts = [merged.x if merged is not None else None]
        self.acts = acts                              # layers read model.acts (multi-scale readout)
        for layer in self._active_layers():
            acts.append(layer(acts[-1], batch_data, self))
        return acts[-1]

    def store_layer_output(self, layer, output):
        self.layer_output[layer] = output

    def get_layer_output(self, layer):
        return self.layer_output[layer]
