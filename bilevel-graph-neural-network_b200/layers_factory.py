"""Layer registry with the reference's spec-string grammar, constructor contracts and
error behaviour (model/layers_factory.py:14-205)."""
import torch.nn as nn

from .config import get_flags
from .layers import NodeEmbedding, Loss
from .layers_aggregation import NodeAggregation, NodeAggregationPairs
from .layers_link_pred import LinkPred
from .layers_load_interaction_graph import LoadInteractionGraph
from .layers_meta import MetaLayerWrapper


def create_layers(model, pattern, num_layers):
    flags = vars(get_flags())
    layers = nn.ModuleList()
    for i in range(1, num_layers + 1):
        parts = flags['{}_{}'.format(pattern, i)].split(':')
        name = parts[0]
        info = {}
        if len(parts) > 1:
            assert len(parts) == 2
            for item in parts[1].split(','):
                kv = item.split('=')
                info[kv[0]] = '='.join(kv[1:])
        if name not in layer_ctors:
            raise ValueError('Unknown layer {}'.format(name))
        layers.append(layer_ctors[name](info, model, i, layers, num_layers))
    return layers


def _check_spec(allowed_nums, lf, ln):
    if len(lf) not in allowed_nums:
        raise ValueError('{} layer must have {} specs NOT {} {}'.format(ln, allowed_nums, len(lf), lf))


def _parse_as_bool(b):
    if b == 'True':
        return True
    if b == 'False':
        return False
    raise RuntimeError('Unknown bool string {}'.format(b))


def get_input_dim_higher_level(lf, lyr_class, layers, layer_id, model):
    input_dim = lf.get('input_dim')
    higher_level = lf.get('higher_level')
    higher_level = _parse_as_bool(higher_level) if higher_level else False
    if input_dim is None:
        if lyr_class in [type(l) for l in layers]:
            raise RuntimeError('The input dim for layer must be specified'.format(layer_id))
        input_dim = model.interaction_num_node_feat if higher_level else model.num_node_feat
    else:
        input_dim = int(input_dim)
    return input_dim, higher_level


def create_node_embedding_layer(lf, model, layer_id, layers, *unused):
    _check_spec([4, 5, 6, 7], lf, 'NodeEmbedding')
    input_dim, higher_level = get_input_dim_higher_level(lf, NodeEmbedding, layers, layer_id, model)
    return NodeEmbedding(type=lf['type'], in_dim=input_dim, out_dim=int(lf['output_dim']), act=lf['act'],
                         bn=_parse_as_bool(lf['bn']), normalize=_parse_as_bool(lf['normalize']),
                         higher_level=higher_level)


def create_meta_wrapper_layer(lf, model, layer_id, layers, *unused):
    _check_spec([5, 6], lf, 'MetaLayerWrapper')
    input_dim, higher_level = get_input_dim_higher_level(lf, MetaLayerWrapper, layers, layer_id, model)
    return MetaLayerWrapper(input_dim=input_dim, edge_dim=model.num_hyper_edge_feat, output_dim=int(lf['output_dim']),
                            edge_model=lf['edge_model'], node_model=lf['node_model'], higher_level=higher_level,
                            num_edge_types=model.num_hyper_edge_feat, act=lf['act'])


def create_load_interaction_graph_layer(lf, *unused):
    _check_spec([0], lf, 'LoadInteractionGraph')
    return LoadInteractionGraph()


def _opt_int(lf, k):
    v = lf.get(k)
    return int(v) if v is not None else None


def create_node_aggregation_layer(lf, model, layer_id, layers, num_layers, *unused):
    _check_spec([1, 3, 4], lf, 'NodeAggregation')
    cms = lf.get('concat_multi_scale')
    return NodeAggregation(style=lf['style'], is_last_layer=(layer_id + 2) == num_layers,
                           concat_multi_scale=_parse_as_bool(cms) if cms is not None else False,
                           in_dim=_opt_int(lf, 'in_dim'), out_dim=_opt_int(lf, 'out_dim'),
                           num_mlp_layers=_opt_int(lf, 'num_mlp_layers'))


def create_node_aggregation_pairs_layer(lf, *unused):
    _check_spec([1], lf, 'NodeAggregationPairs')
    return NodeAggregationPairs(style=lf['style'])


def create_link_pred_layer(lf, model, *unused):
    _check_spec([3, 4, 5], lf, 'LinkPred')
    weight_dim = lf.get('weight_dim')
    weight_dim = int(weight_dim) if weight_dim else weight_dim
    mlp_dim = lf.get('mlp_dim')
    mlp_dim = int(mlp_dim) if mlp_dim else mlp_dim
    multi = lf['multi_label_pred']
    multi = _parse_as_bool(multi) if multi else False
    return LinkPred(type=lf['type'], mlp_dim=mlp_dim, weight_dim=weight_dim,
                    batch_unique_graphs=_parse_as_bool(lf['batch_unique_graphs']),
                    multi_label_pred=multi, num_labels=model.num_labels + 1)


def create_loss_layer(lf, *unused):
    _check_spec([1], lf, 'Loss')
    return Loss(type=lf['type'])


def _later(name):
    def ctor(*unused):
        raise NotImplementedError('{} is outside the Bi-GNN hot path built so far'.format(name))
    return ctor


layer_ctors = {
    'NodeEmbedding': create_node_embedding_layer,
    'NodeAggregation': create_node_aggregation_layer,
    'NodeAggregationPairs': create_node_aggregation_pairs_layer,
    'Loss': create_loss_layer,
    'GMNPropagator': _later('GMNPropagator'),
    'GMNAggregatorPairs': _later('GMNAggregatorPairs'),
    'LinkPredictor': create_link_pred_layer,
    'MetaLayer': create_meta_wrapper_layer,
    'LoadInteractionLayer': create_load_interaction_graph_layer,
}
