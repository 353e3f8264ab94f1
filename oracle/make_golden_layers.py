"""TEST INFRASTRUCTURE ONLY -- layer-level golden vectors for the NON-DEFAULT variants of the layers on the path,
recorded from the reference's own layer classes (model/layers.py, model/layers_aggregation.py,
model/layers_link_pred.py) on one merged batch of 10 DrugBank molecules: NodeEmbedding x {gin, gcn, gat} x
{relu, prelu, sigmoid, tanh, identity} x bn x normalize, the four readout styles (avg_pool, sum, deepsets, gmn_aggr;
single- and multi-scale), the MLP scorer and the BCE / BCEWithLogits losses.  For each case: the module's state_dict, the
output, and the gradients of sum(output * R) w.r.t. every parameter and the input (R fixed).
-> tests/golden/bignn_layer_variants.npz.  Runs only where /root/reference exists.

Usage:  python oracle/make_golden_layers.py [--out tests/golden]"""
import argparse
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    args = ap.parse_args()
    # model='lower_level_gnn': FLAGS.higher_level_layers is False, so NodeAggregation returns the pooled rows
    # without touching an interaction graph (model/layers_aggregation.py:68-69)
    ref_loader.load_reference(model='lower_level_gnn')
    train_data, _, _, FLAGS = ref_loader.load_drugbank_fold(1)
    import torch
    from batch import BatchData
    from model.layers import NodeEmbedding, Loss
    from model.layers_aggregation import NodeAggregation
    from model.layers_link_pred import LinkPred
    ds = train_data.dataset
    gids = list(ds.gs_map.keys())[100:110]
    pairs = np.asarray([(gids[i], gids[i + 1]) for i in range(0, 10, 2)], np.int64)
    bd = BatchData(pairs, ds, is_train=False, ignore_pairs=True)
    merge = bd.merge_data['merge']
    G = len(bd.merge_data['gids_to_batch_ind'])
    A = merge.x.shape[0]
    out = dict(gids=np.asarray(list(bd.merge_data['gids_to_batch_ind'].keys()), np.int64), pairs=pairs,
               edge_index=merge.edge_index.numpy().astype(np.int32), batch=merge.batch.numpy().astype(np.int32),
               x_u8=merge.x.numpy().astype(np.uint8))
    rng = torch.Generator().manual_seed(1234)
    h64 = torch.randn(A, 64, generator=rng)
    R_nodes = torch.randn(A, 64, generator=rng)
    R_graphs = torch.randn(G, 320, generator=rng)
    out['h64'], out['R_nodes'], out['R_graphs'] = h64.numpy(), R_nodes.numpy(), R_graphs.numpy()
    model = types.SimpleNamespace(acts=None, store_layer_output=lambda layer, x: None)
    cases = []

    def record(tag, module, fn, x_in):
        x = x_in.clone().requires_grad_(True)
        module.train()
        sd = {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}
        y, R = fn(module, x)
        (y * R).sum().backward()
        out[tag + '/out'] = y.detach().numpy()
        out[tag + '/dx'] = x.grad.numpy()
        for k, v in sd.items():
            out[tag + '/sd/' + k] = v
        for k, p in module.named_parameters():
            out[tag + '/grad/' + k] = p.grad.numpy() if p.grad is not None else np.zeros_like(p.detach().numpy())
        for k, v in module.state_dict().items():
            if 'running' in k:
                out[tag + '/sd1/' + k] = v.detach().numpy().copy()
        cases.append(tag)

    # ---- NodeEmbedding variants (model/layers.py:9-63)
    for typ in ('gin', 'gcn', 'gat'):
        for act in ('relu', 'prelu', 'sigmoid', 'tanh', 'identity'):
            for bn, norm in ((True, False), (False, True), (True, True)):
                plain = bn and not norm
                if typ != 'gin' and not ((act in ('relu', 'prelu', 'tanh') and plain) or act == 'relu'):
                    continue                                     # keep the fixture small (GIN gets the full grid)
                torch.manual_seed(7)
                m = NodeEmbedding(typ, 64, 64, act, bn, norm)
                tag = 'ne/{}/{}/bn{}/norm{}'.format(typ, act, int(bn), int(norm))
                record(tag, m, lambda mod, x: (mod(x, bd, model), R_nodes), h64)
    # first-layer shape (one-hot input, F_in columns)
    torch.manual_seed(7)
    m = NodeEmbedding('gin', merge.x.shape[1], 64, 'relu', True, False)
    record('ne/gin/first_layer', m, lambda mod, x: (mod(x, bd, model), R_nodes), merge.x.float())
    # ---- readouts (model/layers_aggregation.py)
    acts5 = [torch.randn(A, 64, generator=rng) for _ in range(5)]
    for i, a in enumerate(acts5):
        out['acts5/%d' % i] = a.numpy()
    for style in ('avg_pool', 'sum'):
        torch.manual_seed(7)
        m = NodeAggregation(style, True, concat_multi_scale=True, in_dim=64, out_dim=64)

        def multi(mod, x):
            model.acts = [None] + [x] + acts5[1:]
            return mod(x, bd, model), R_graphs
        record('agg/{}/multi'.format(style), m, multi, acts5[0])
        torch.manual_seed(7)
        m = NodeAggregation(style, True, concat_multi_scale=False, in_dim=64, out_dim=64)
        record('agg/{}/single'.format(style), m, lambda mod, x: (mod(x, bd, model), R_graphs[:, :64]), acts5[0])
    torch.manual_seed(7)
    m = NodeAggregation('deepsets', True, concat_multi_scale=False, in_dim=64, out_dim=64, num_mlp_layers=2)
    record('agg/deepsets', m, lambda mod, x: (mod(x, bd, model), R_graphs[:, :64]), acts5[0])
    torch.manual_seed(7)
    m = NodeAggregation('gmn_aggr', True, concat_multi_scale=False, in_dim=64, out_dim=64)
    record('agg/gmn_aggr', m, lambda mod, x: (mod(x, bd, model), R_graphs[:, :64]), acts5[0])
    # ---- scorers and losses on graph embeddings [G, 64] (model/layers_link_pred.py, model/layers.py:66-89)
    pairs2 = train_data.data_items[:12].numpy().astype(np.int64)     # real train pairs (labels exist)
    bd2 = BatchData(pairs2, ds, is_train=False)
    out['pairs2'] = pairs2
    out['pairs2_gids'] = np.asarray(list(bd2.merge_data['gids_to_batch_ind'].keys()), np.int64)
    emb = torch.randn(len(out['pairs2_gids']), 64, generator=rng)
    out['emb'] = emb.numpy()
    Rp = torch.randn(len(pairs2), generator=rng)
    out['R_pairs'] = Rp.numpy()
    # ('dot_product' cannot be recorded: the reference's own path raises IndexError at utils/data/graph.py:36 -- its
    # scores are 0-d tensors and assign_link_pred indexes .shape[0])
    for typ in ('mlp_concat',):
        torch.manual_seed(7)
        m = LinkPred(typ, 64, 2)
        record('lp/' + typ, m, lambda mod, x: (mod(x, bd2, model).view(-1), Rp), emb)
    out['y_pairs'] = np.asarray([p.true_label for p in bd2.pair_list], np.int64)
    logits = torch.randn(len(pairs2), generator=rng)
    out['logits'] = logits.numpy()
    for typ in ('BCE', 'BCEWithLogits'):
        m = Loss(typ)
        x = (torch.sigmoid(logits) if typ == 'BCE' else logits).clone().requires_grad_(True)
        l = m(x.view(-1, 1), bd2, model)
        l.backward()
        out['loss/' + typ + '/out'] = np.float32(l.item())
        out['loss/' + typ + '/dx'] = x.grad.numpy()
    out['cases'] = np.asarray(cases)
    path = os.path.join(args.out, 'bignn_layer_variants.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(cases), 'cases', 'atoms', A, 'graphs', G)


if __name__ == '__main__':
    main()
