"""The architecture the reference ships for DrugCombo -- GIN x5 lower, multi-scale mean readout, 3 x MetaLayer (one
GAT per interaction edge type, summed: model/layers_meta.py:61-79), 3-class MLP scorer, cross entropy -- on the GPU
against vectors recorded from the REFERENCE'S OWN CODE on the DrugCombo subset its tree still holds
(tests/golden/bignn_drugcombo_step.npz, oracle/make_golden.py --dataset drugcombo; 1 621 drugs, synergy and
antagonism interaction graphs).  Same gates as tests/test_gpu_step.py: forward 1e-5, gradients no further from the
fp64 oracle than 6x the reference's own fp32 gradients are (+2e-5).
Measured on B200 (profiles/r1b_golden_gpu.log): init_x 6.9e-6, MetaLayer activations 3.0e-6, logits 4.6e-6, loss
equal to 7 digits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200.engine import BiGNNEngine, _StaticPairBatch
from oracle import bignn_oracle as O

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def golden(golden_dir):
    z = np.load(os.path.join(golden_dir, 'bignn_drugcombo_step.npz'))
    with open(os.path.join(golden_dir, 'bignn_drugcombo_layers.txt')) as f:
        lines = f.read().split()
    return z, lines


def build(golden_dir, z, lines):
    B._lib.load()
    flags = B.make_flags(dataset='drugcombo', higher_level_gnn_type='gat', device=DEV)
    B.set_flags(flags)
    assert [getattr(flags, 'layer_%d' % i) for i in range(1, flags.layer_num + 1)] == lines
    data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugcombo_packed.npz'), device=DEV)
    assert list(data.interaction_nxgraphs) == ['0_synergy', '1_antagonism']
    model = B.Model(data).to(DEV)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    for k in z.files:
        if k.startswith('sd_init/'):
            sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert {k for k in model.state_dict() if k.startswith('layers.')} == {k[4:] for k in z.files if k.startswith('sd0/')}
    model.train()
    return data, model


def test_drugcombo_step_vs_reference_golden(golden_dir, golden):
    z, lines = golden
    try:
        data, model = build(golden_dir, z, lines)
        eng = BiGNNEngine(data, model, use_cuda_graph=False)
        assert eng.n_chunks_total == int(z['n_chunks'])
        st, P = eng.stage_pairs(z['batch_gids'], z['y_true'].astype(np.float32))
        sb = _StaticPairBatch(data, P, data.device)
        sb.load(st)
        loss = eng.forward(sb)
        loss.backward()
        acts = model.acts                   # [None, LoadInteraction, MetaLayer x3, LinkPred, Loss]
        errs = dict(init_x=rel(data.interaction_combo_nxgraph.init_x, z['init_x']), act2=rel(acts[2], z['upper/act2']),
                    act4=rel(acts[4], z['upper/act4']), logits=rel(sb.preds, z['upper/act5']),
                    loss=abs(float(loss.detach()) - float(z['loss'])))
        print('drugcombo golden, forward errors:', {k: float('%.3g' % v) for k, v in errs.items()})
        assert errs['init_x'] < 1e-5 and errs['act2'] < 1e-5 and errs['act4'] < 1e-5
        assert errs['logits'] < 1e-5          # [128, 3] logits
        assert errs['loss'] < 1e-5
        # gradients: fp64 ground truth from the oracle (itself pinned to this golden, tests/test_oracle_golden.py)
        ds = O.PackedDataset.load(os.path.join(golden_dir, 'drugcombo_packed.npz'))
        om = O.OracleModel(O.parse_specs(lines), O.state_from_npz(z, 'sd0/'), dtype=torch.float64, gat_group='source')
        _, _, _, l64 = O.train_step_forward(om, ds, z['batch_gids'], z['y_true'])
        l64.backward()
        g64 = {k: v.grad.numpy() for k, v in om.params().items() if v.grad is not None}
        scale = {}
        for k, g in g64.items():
            scale[k.split('.')[1]] = max(scale.get(k.split('.')[1], 0.0), float(np.abs(g).max()))
        named = dict(model.named_parameters())
        n = 0
        for k, g in g64.items():
            s = scale[k.split('.')[1]]
            ours = float(np.abs(named[k].grad.double().cpu().numpy() - g).max()) / s
            ref = float(np.abs(z['grad/' + k].astype(np.float64) - g).max()) / s
            assert ours <= 6.0 * ref + 2e-5, (k, ours, ref)
            n += 1
        assert n == len([k for k in z.files if k.startswith('grad/')])
    finally:
        B.set_flags(B.make_flags(device=DEV))


def test_drugcombo_reference_sequenced_path_and_bn_buffers(golden_dir, golden):
    """the layer-by-layer driver (train.py, chunk by chunk as src/train.py:48-72) on the same step: loss and the
    BatchNorm running buffers after the 13 sequential chunk updates"""
    z, lines = golden
    try:
        data, model = build(golden_dir, z, lines)
        model.zero_grad()
        B.train._get_initial_embd(data, model)
        bd = B.BatchData(z['positive_gids'], data, sampled_gids=z['sampled_gids'], is_train=False, merge_graphs=False)
        bd.batch_gids = z['batch_gids']
        bd.pair_list = [B.batch.PairRecord(int(l), tuple(g)) for l, g in zip(z['y_true'], z['batch_gids'].tolist())]
        bd.batch_interaction_inds = [data.gs_map[g] for g in bd.batch_gids.flatten().tolist()]
        model.use_layers = 'higher_layers'
        loss = model(bd)
        assert abs(float(loss.detach()) - float(z['loss'])) < 1e-5
        assert rel(data.interaction_combo_nxgraph.init_x, z['init_x']) < 1e-5
        sdm = model.state_dict()
        for k in z.files:
            if k.startswith('sd1/') and 'running' in k:
                assert rel(sdm[k[4:]], z[k]) < 1e-5, k
            if k.startswith('sd1/') and 'num_batches' in k:
                assert int(sdm[k[4:]]) == int(z[k]), k
    finally:
        B.set_flags(B.make_flags(device=DEV))
