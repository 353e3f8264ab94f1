"""Run configuration of the path.

The reference keeps its configuration in an argparse namespace built at import time
(src/config.py:467 `FLAGS`) whose `layer_<i>` strings are the input of the layer
registry (model/layers_factory.py:16-24).  `make_flags` produces the same namespace
fields -- and, for a given (model, gnn types, depths), the same `layer_<i>` strings
the reference generates (src/config.py:297-433) -- without argv / import-time side
effects.  When this package is used inside the reference tree, `get_flags()` returns
the reference's own `config.FLAGS` instead.
"""
import sys
from types import SimpleNamespace

_FLAGS = None


def make_flags(model='lower_level_gnn_higher_level', lower_level_gnn_type='gin', higher_level_gnn_type='gcn',
               lower_level_num_layers=5, higher_level_num_layers=3, D_lower=64, D_higher=64,
               node_aggr='multi_scale', style='avg_pool', bn=True, gnn_normalize=False,
               multi_class_pred=None, batch_size=64, device='cuda:0', dataset='drugbank', **extra):
    # DrugCombo defaults (src/config.py:74-88,120-123,224): one GNN per interaction edge type through
    # MetaLayer, 3-class prediction with cross entropy
    combo = 'drugcombo' in dataset
    if multi_class_pred is None:
        multi_class_pred = combo
    node_model = extra.pop('node_model', '{}_multi_edge_aggr'.format(higher_level_gnn_type) if combo else None)
    lower = 'lower_level' in model
    higher = 'higher_level' in model
    specs = []
    f = SimpleNamespace()
    d_lower = D_lower
    if lower:
        g = lower_level_gnn_type
        specs.append('NodeEmbedding:type={},output_dim={},act=relu,bn={},normalize={}'.format(
            g, d_lower, bn, gnn_normalize))
        for i in range(lower_level_num_layers - 1):
            act = 'identity' if i == lower_level_num_layers - 2 else 'relu'
            specs.append('NodeEmbedding:type={},input_dim={},output_dim={},act={},bn={},normalize={}'.format(
                g, d_lower, d_lower, act, bn, gnn_normalize))
        if node_aggr == 'multi_scale':
            specs.append('NodeAggregation:style={},concat_multi_scale={},in_dim={},out_dim={}'.format(
                style, True, d_lower, d_lower))
            d_lower = d_lower * lower_level_num_layers
        elif node_aggr in ('avg_pool', 'sum'):
            specs.append('NodeAggregation:style={}'.format(node_aggr))
        else:
            raise NotImplementedError(node_aggr)
    init_embds = 'model_init' if (lower and higher) else ('no_init' if lower else 'rand_init')
    if lower and 'model_init' in init_embds:
        f.last_lower_lyr_num = len(specs) - 1
    if higher:
        specs.append('LoadInteractionLayer')
        g = higher_level_gnn_type
        n_h = higher_level_num_layers
        if combo:
            if lower:
                specs.append('MetaLayer:input_dim={},output_dim={},act=relu,higher_level={},edge_model={},'
                             'node_model={}'.format(d_lower, D_higher, True, 'none', node_model))
            else:
                specs.append('MetaLayer:output_dim={},act=relu,higher_level={},edge_model={},node_model={}'.format(
                    D_higher, True, 'none', node_model))
            for i in range(n_h - 1):
                act = 'identity' if i == n_h - 2 else 'relu'
                specs.append('MetaLayer:input_dim={},output_dim={},act={},higher_level={},edge_model={},'
                             'node_model={}'.format(D_higher, D_higher, act, True, 'none', node_model))
        elif lower:
            specs.append('NodeEmbedding:type={},input_dim={},output_dim={},act=relu,bn={},higher_level={},'
                         'normalize={}'.format(g, d_lower, D_higher, bn, True, gnn_normalize))
        else:
            specs.append('NodeEmbedding:type={},output_dim={},act=relu,bn={},higher_level={},normalize={}'.format(
                g, D_higher, bn, True, gnn_normalize))
        for i in range(n_h - 1 if not combo else 0):
            act = 'identity' if i == n_h - 2 else 'relu'
            specs.append('NodeEmbedding:type={},input_dim={},output_dim={},act={},bn={},higher_level={},'
                         'normalize={}'.format(g, D_higher, D_higher, act, bn, True, gnn_normalize))
    else:
        D_higher = d_lower
    specs.append('LinkPredictor:type=mlp_concat,multi_label_pred={},mlp_dim={},batch_unique_graphs={}'.format(
        multi_class_pred, D_higher, True))
    specs.append('Loss:type={}'.format('CE' if multi_class_pred else 'BCE'))
    for i, s in enumerate(specs):
        setattr(f, 'layer_%d' % (i + 1), s)
    f.layer_num = len(specs)
    f.model = model
    f.dataset = dataset
    f.lower_level_layers = lower
    f.higher_level_layers = higher
    f.init_embds = init_embds
    f.pair_interaction = False
    f.batch_size = batch_size
    f.batch_unique_graphs = True
    f.device = device
    f.lr = 1e-3
    f.random_seed = 3
    f.negative_sample = True
    f.num_negative_samples = 1
    f.enforce_negative_sampling = True
    f.enforce_sampling_amongst_same_graphs = True
    f.sample_induced = False
    f.different_edge_type_aggr = combo
    f.use_hyper_edge_attrs = False
    f.multi_class_pred = multi_class_pred
    f.d_init = 64
    for k, v in extra.items():
        setattr(f, k, v)
    return f


def set_flags(flags):
    global _FLAGS
    _FLAGS = flags
    return flags


def get_flags():
    """The active namespace: an explicitly set one, else the reference's `config.FLAGS`
    when this package runs inside the reference tree, else the Bi-GNN defaults."""
    global _FLAGS
    if _FLAGS is not None:
        return _FLAGS
    ref = sys.modules.get('config')
    if ref is not None and hasattr(ref, 'FLAGS'):
        return ref.FLAGS
    _FLAGS = make_flags()
    return _FLAGS
