"""Layer registry of the path.

Same contract as the reference's registry (model/layers_factory.py:14-205): layers are declared by
`Name:key=value,key=value` strings in `FLAGS.<pattern>_<i>`, `layer_ctors[name]` is called as
`ctor(spec_dict, model, layer_id (1-based), layers_built_so_far, num_layers)`, a wrong number of keys
is a ValueError, bools are the literal strings 'True' / 'False' (anything else: RuntimeError), an
unknown layer name is a ValueError.  The layer classes behind it are this package's CUDA-backed ones.
"""
import torch.nn as nn

from .config import get_flags
from .layers import NodeEmbedding, Loss
from .layers_aggregation import NodeAggregation, NodeAggregationPairs
from .layers_link_pred import LinkPred
from .layers_load_interaction_graph import LoadInteractionGraph
from .layers_meta import MetaLayerWrapper


# ----------------------------------------------------------------------------- spec handling
def parse_layer_spec(text):
    """'Name:k=v,k=v' -> (name, {k: v}); values may themselves contain '='."""
    name, _, body = text.partition(':')
    if ':' in body:
        raise AssertionError('a layer spec has at most one ":" ({!r})'.format(text))
    info = {}
    if body:
        for item in body.split(','):
            key, _, value = item.partition('=')
            info[key] = value
    return name, info


def _check_spec(allowed_nums, lf, ln):
    if len(lf) not in allowed_nums:
        raise ValueError('{} layer must have {} specs NOT {} {}'.format(ln, allowed_nums, len(lf), lf))


def _parse_as_bool(b):
    if b in ('True', 'False'):
        return b == 'True'
    raise RuntimeError('Unknown bool string {}'.format(b))


def _maybe(lf, key, conv, default=None):
    v = lf.get(key)
    return conv(v) if v not in (None, '') else default


def create_layers(model, pattern, num_layers):
    flags = vars(get_flags())
    built = nn.ModuleList()
    for layer_id in range(1, num_layers + 1):
        name, info = parse_layer_spec(flags['{}_{}'.format(pattern, layer_id)])
        ctor = layer_ctors.get(name)
        if ctor is None:
            raise ValueError('Unknown layer {}'.format(name))
        built.append(ctor(info, model, layer_id, built, num_layers))
    return built


def get_input_dim_higher_level(lf, lyr_class, layers, layer_id, model):
    """input_dim given in the spec, else the dataset's feature width for the FIRST layer of that class
    (molecule features downstairs, interaction-graph features upstairs)."""
    higher_level = _maybe(lf, 'higher_level', _parse_as_bool, False)
    input_dim = _maybe(lf, 'input_dim', int)
    if input_dim is None:
        if any(type(l) is lyr_class for l in layers):
            raise RuntimeError('The input dim for layer {} must be specified'.format(layer_id))
        input_dim = model.interaction_num_node_feat if higher_level else model.num_node_feat
    return input_dim, higher_level


# ----------------------------------------------------------------------------- constructors
def create_node_embedding_layer(lf, model, layer_id, layers, *unused):
    _check_spec([4, 5, 6, 7], lf, 'NodeEmbedding')
    in_dim, higher = get_input_dim_higher_level(lf, NodeEmbedding, layers, layer_id, model)
    return NodeEmbedding(type=lf['type'], in_dim=in_dim, out_dim=int(lf['output_dim']), act=lf['act'],
                         bn=_parse_as_bool(lf['bn']), normalize=_parse_as_bool(lf['normalize']), higher_level=higher)


def create_meta_wrapper_layer(lf, model, layer_id, layers, *unused):
    _check_spec([5, 6], lf, 'MetaLayerWrapper')
    in_dim, higher = get_input_dim_higher_level(lf, MetaLayerWrapper, layers, layer_id, model)
    n_types = model.num_hyper_edge_feat
    return MetaLayerWrapper(input_dim=in_dim, edge_dim=n_types, output_dim=int(lf['output_dim']),
                            edge_model=lf['edge_model'], node_model=lf['node_model'], act=lf['act'],
                            num_edge_types=n_types, higher_level=higher)


def create_load_interaction_graph_layer(lf, *unused):
    _check_spec([0], lf, 'LoadInteractionGraph')
    return LoadInteractionGraph()


def create_node_aggregation_layer(lf, model, layer_id, layers, num_layers, *unused):
    _check_spec([1, 3, 4], lf, 'NodeAggregation')
    return NodeAggregation(style=lf['style'],
                           is_last_layer=(layer_id + 2 == num_layers),      # followed by scorer and loss only
                           concat_multi_scale=_maybe(lf, 'concat_multi_scale', _parse_as_bool, False),
                           in_dim=_maybe(lf, 'in_dim', int), out_dim=_maybe(lf, 'out_dim', int),
                           num_mlp_layers=_maybe(lf, 'num_mlp_layers', int))


def create_node_aggregation_pairs_layer(lf, *unused):
    _check_spec([1], lf, 'NodeAggregationPairs')
    return NodeAggregationPairs(style=lf['style'])


def create_link_pred_layer(lf, model, *unused):
    _check_spec([3, 4, 5], lf, 'LinkPred')
    return LinkPred(type=lf['type'], mlp_dim=_maybe(lf, 'mlp_dim', int), weight_dim=_maybe(lf, 'weight_dim', int),
                    batch_unique_graphs=_parse_as_bool(lf['batch_unique_graphs']),
                    multi_label_pred=_maybe(lf, 'multi_label_pred', _parse_as_bool, False),
                    num_labels=model.num_labels + 1)


def create_loss_layer(lf, *unused):
    _check_spec([1], lf, 'Loss')
    return Loss(type=lf['type'])


def _outside_path(name):
    def ctor(*unused):
        raise NotImplementedError('{} belongs to the MHCADDI baseline, outside the Bi-GNN hot path'.format(name))
    return ctor


layer_ctors = {
    'NodeEmbedding': create_node_embedding_layer,
    'NodeAggregation': create_node_aggregation_layer,
    'NodeAggregationPairs': create_node_aggregation_pairs_layer,
    'Loss': create_loss_layer,
    'GMNPropagator': _outside_path('GMNPropagator'),
    'GMNAggregatorPairs': _outside_path('GMNAggregatorPairs'),
    'LinkPredictor': create_link_pred_layer,
    'MetaLayer': create_meta_wrapper_layer,
    'LoadInteractionLayer': create_load_interaction_graph_layer,
}
