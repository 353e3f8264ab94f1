"""The oracle's restatement of the NON-DEFAULT layer variants on the path -- NodeEmbedding {gin, gcn, gat} x
{relu, prelu, sigmoid, tanh, identity} x BatchNorm x normalize, the sum / deepsets / gated ("gmn_aggr") readouts,
BCEWithLogits -- pinned against the reference's own layer classes (oracle/make_golden_layers.py ->
tests/golden/bignn_layer_variants.npz): outputs, input gradients, every parameter gradient, BatchNorm running
buffers.  (The GPU kernels are compared with the oracle on these variants in tests/test_gpu_more_layers.py.)"""
import os

import numpy as np
import pytest
import torch

from oracle import bignn_oracle as O


def rel(a, b, floor=1e-30):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


@pytest.fixture(scope='module')
def lv(golden_dir):
    torch.set_num_threads(8)
    return np.load(os.path.join(golden_dir, 'bignn_layer_variants.npz'))


def _state(z, tag, prefix='l.'):
    n = len(tag) + 4
    return {prefix + k[n:]: torch.from_numpy(np.asarray(z[k])).clone() for k in z.files if k.startswith(tag + '/sd/')}


def _check_grads(z, tag, P, prefix='l.', tol=2e-5):
    n = 0
    scale = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith(tag + '/grad/'))
    for k in z.files:
        if k.startswith(tag + '/grad/'):
            name = prefix + k[len(tag) + 6:]
            g = P[name].grad
            got = np.zeros_like(z[k]) if g is None else g.numpy()
            assert np.abs(got.astype(np.float64) - z[k]).max() <= tol * max(scale, 1e-12), (tag, name)
            n += 1
    return n


def test_node_embedding_variants_match_reference(lv):
    z = lv
    ei = torch.from_numpy(z['edge_index'].astype(np.int64))
    cases = [c for c in z['cases'].tolist() if c.startswith('ne/')]
    assert len(cases) >= 20
    for tag in cases:
        parts = tag.split('/')
        if parts[2] == 'first_layer':
            lf = dict(type='gin', act='relu', bn='True', normalize='False')
            x = torch.from_numpy(z['x_u8'].astype(np.float32))
        else:
            lf = dict(type=parts[1], act=parts[2], bn=str(parts[3] == 'bn1'), normalize=str(parts[4] == 'norm1'))
            x = torch.from_numpy(z['h64'])
        P = _state(z, tag)
        for k, v in P.items():
            if v.dtype.is_floating_point and 'running' not in k and not k.endswith('conv.eps'):
                v.requires_grad_(True)
        x = x.clone().requires_grad_(True)
        out = O.node_embedding(x, ei, P, 'l', lf, True, 'source')
        (out * torch.from_numpy(z['R_nodes'])).sum().backward()
        assert rel(out.detach().numpy(), z[tag + '/out']) < 2e-6, tag
        assert rel(x.grad.numpy(), z[tag + '/dx'], 1e-12) < 5e-5, tag
        assert _check_grads(z, tag, P) >= 2
        for k in z.files:
            if k.startswith(tag + '/sd1/'):
                assert rel(P['l.' + k[len(tag) + 5:]].detach().numpy(), z[k]) < 1e-6, (tag, k)


def test_readout_variants_match_reference(lv):
    z = lv
    batch = torch.from_numpy(z['batch'].astype(np.int64))
    G = int(z['gids'].shape[0])
    acts5 = [torch.from_numpy(z['acts5/%d' % i]) for i in range(5)]
    Rg = torch.from_numpy(z['R_graphs'])
    for style in ('avg_pool', 'sum'):
        x = acts5[0].clone().requires_grad_(True)
        out = O.readout([x] + acts5[1:], batch, G, style)
        (out * Rg).sum().backward()
        assert rel(out.detach().numpy(), z['agg/%s/multi/out' % style]) < 1e-6
        assert rel(x.grad.numpy(), z['agg/%s/multi/dx' % style]) < 1e-6
        x = acts5[0].clone().requires_grad_(True)
        out = O.readout([x], batch, G, style)
        (out * Rg[:, :64]).sum().backward()
        assert rel(out.detach().numpy(), z['agg/%s/single/out' % style]) < 1e-6
        assert rel(x.grad.numpy(), z['agg/%s/single/dx' % style]) < 1e-6
    for tag, fn in (('agg/deepsets', lambda x, P: O.deepsets_readout(x, batch, G, P, 'l.agg_func', 2)),
                    ('agg/gmn_aggr', lambda x, P: O.gmn_aggr_readout(x, batch, G, P, 'l.agg_func'))):
        P = _state(z, tag)
        for k, v in P.items():
            if v.dtype.is_floating_point and 'running' not in k:
                v.requires_grad_(True)
        x = acts5[0].clone().requires_grad_(True)
        out = fn(x, P)
        (out * Rg[:, :64]).sum().backward()
        assert rel(out.detach().numpy(), z[tag + '/out']) < 2e-6, tag
        assert rel(x.grad.numpy(), z[tag + '/dx'], 1e-12) < 5e-5, tag
        assert _check_grads(z, tag, P, tol=5e-5) >= 4


def test_scorer_and_losses_match_reference(lv):
    z = lv
    order = {int(g): i for i, g in enumerate(z['pairs2_gids'].tolist())}
    ids = torch.from_numpy(np.vectorize(order.get)(z['pairs2']).astype(np.int64))
    P = _state(z, 'lp/mlp_concat')
    for v in P.values():
        v.requires_grad_(True)
    emb = torch.from_numpy(z['emb']).clone().requires_grad_(True)
    pred = O.link_pred(emb, ids, P, 'l', dict(type='mlp_concat'), 2)
    (pred.view(-1) * torch.from_numpy(z['R_pairs'])).sum().backward()
    assert rel(pred.detach().numpy().reshape(-1), z['lp/mlp_concat/out']) < 1e-6
    assert rel(emb.grad.numpy(), z['lp/mlp_concat/dx'], 1e-12) < 2e-5
    assert _check_grads(z, 'lp/mlp_concat', P) == 6
    y = torch.from_numpy(z['y_pairs'])
    logits = torch.from_numpy(z['logits'])
    for kind in ('BCE', 'BCEWithLogits'):
        x = (torch.sigmoid(logits) if kind == 'BCE' else logits).clone().requires_grad_(True)
        l = O.loss_fn(x.view(-1, 1), y, kind)
        l.backward()
        assert abs(float(l.detach()) - float(z['loss/%s/out' % kind])) < 1e-6
        assert rel(x.grad.numpy(), z['loss/%s/dx' % kind]) < 1e-6
