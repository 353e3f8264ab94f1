"""Lower-level-only model (model='lower_level_gnn', the LL-GNN baseline = the model of BASELINE config 3) on the GPU
against the reference's own recorded step (tests/golden/bignn_ll_gnn_step.npz).  The same fixture is green for the
oracle and for the host path on the CPU stand-in backend.  First seen green on a B200 in round 2
(gpurun_out/pytest_r2a.log: 146 passed)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]

import bignn_b200 as B
from oracle import bignn_oracle as O

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_ll_gnn_step_vs_reference_golden(golden_dir, drugbank):
    B._lib.load()
    z = np.load(os.path.join(golden_dir, 'bignn_ll_gnn_step.npz'))
    with open(os.path.join(golden_dir, 'bignn_ll_gnn_layers.txt')) as f:
        lines = f.read().split()
    try:
        flags = B.make_flags(model='lower_level_gnn', device=DEV)
        B.set_flags(flags)
        assert [getattr(flags, 'layer_%d' % i) for i in range(1, flags.layer_num + 1)] == lines
        data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
        model = B.Model(data).to(DEV)
        sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected
        model.train()
        model.zero_grad()
        bd = B.BatchData(z['batch_gids'], data, is_train=False)
        m = bd.merge_data['merge']
        assert np.array_equal(m.edge_index.cpu().numpy(), z['edge_index'].astype(np.int64))       # bit-exact
        assert np.array_equal(m.batch.cpu().numpy(), z['batch'].astype(np.int64))
        loss = model(bd)
        errs = dict(act1=rel(model.acts[1], z['act1']), act5=rel(model.acts[5], z['act5']),
                    pooled=rel(model.acts[6], z['act6']), preds=rel(model.acts[7].view(-1), z['act7'].reshape(-1)),
                    loss=abs(float(loss.detach()) - float(z['loss'])))
        print('ll-gnn golden, forward errors:', {k: float('%.3g' % v) for k, v in errs.items()})
        assert max(errs.values()) < 1e-5, errs
        loss.backward()
        om = O.OracleModel(O.parse_specs(lines), O.state_from_npz(z, 'sd0/'), dtype=torch.float64)
        _, _, _, _, l64 = O.lower_only_step_forward(om, drugbank, z['batch_gids'], z['y_true'])
        l64.backward()
        g64 = {k: v.grad.numpy() for k, v in om.params().items()}
        scale = {}
        for k, g in g64.items():
            scale[k.split('.')[1]] = max(scale.get(k.split('.')[1], 0.0), float(np.abs(g).max()))
        named = dict(model.named_parameters())
        for k, g in g64.items():
            s = scale[k.split('.')[1]]
            ours = float(np.abs(named[k].grad.double().cpu().numpy() - g).max()) / s
            ref = float(np.abs(z['grad/' + k].astype(np.float64) - g).max()) / s
            assert ours <= 6.0 * ref + 2e-5, (k, ours, ref)
        sdm = model.state_dict()
        for k in z.files:
            if k.startswith('sd1/') and 'running' in k:
                assert rel(sdm[k[4:]], z[k]) < 1e-5, k
    finally:
        B.set_flags(B.make_flags(device=DEV))


def test_lower_only_engine_fused_vs_layers_vs_golden(golden_dir, drugbank):
    """engine_lower.LowerOnlyEngine on the recorded LL-GNN step: the fused lower stack (one BatchNorm batch = the whole
    merged pair batch) and the layer-by-layer stack against the reference's loss / pooled rows / gradients; then a
    large pair batch (the engine's reason to exist) with the vectorised negative sampler."""
    from bignn_b200.engine_lower import LowerOnlyEngine
    B._lib.load()
    z = np.load(os.path.join(golden_dir, 'bignn_ll_gnn_step.npz'))
    try:
        B.set_flags(B.make_flags(model='lower_level_gnn', device=DEV))
        data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
        res = {}
        for fused_on in (False, True):
            model = B.Model(data).to(DEV)
            sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
            model.load_state_dict(sd, strict=False)
            model.train()
            eng = LowerOnlyEngine(data, model, fused_lower=fused_on)
            assert eng.lower_path == ('fused' if fused_on else 'layers')
            rows, ids, labels = eng.stage(z['batch_gids'], z['y_true'])
            model.zero_grad()
            loss = eng.forward(rows, ids, labels)
            loss.backward()
            assert abs(float(loss.detach()) - float(z['loss'])) < 1e-5
            assert rel(eng.last['pooled'], z['act6']) < 1e-5
            assert rel(eng.last['scores'].view(-1), z['act7'].reshape(-1)) < 1e-5
            res[fused_on] = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
            sdm = model.state_dict()
            for k in z.files:
                if k.startswith('sd1/') and 'running' in k:
                    assert rel(sdm[k[4:]], z[k]) < 1e-5, k
        om = O.OracleModel(O.parse_specs(open(os.path.join(golden_dir, 'bignn_ll_gnn_layers.txt')).read().split()),
                           O.state_from_npz(z, 'sd0/'), dtype=torch.float64)
        _, _, _, _, l64 = O.lower_only_step_forward(om, drugbank, z['batch_gids'], z['y_true'])
        l64.backward()
        g64 = {k: v.grad.numpy() for k, v in om.params().items()}
        scale = {}
        for k, g in g64.items():
            scale[k.split('.')[1]] = max(scale.get(k.split('.')[1], 0.0), float(np.abs(g).max()))
        for fused_on in (False, True):
            for k, g in g64.items():
                s = scale[k.split('.')[1]]
                ours = float(np.abs(res[fused_on][k].double().cpu().numpy() - g).max()) / s
                ref = float(np.abs(z['grad/' + k].astype(np.float64) - g).max()) / s
                # the fused stack computes the hidden activations on the tensor cores (3xTF32) where the layer path
                # uses fp32 FMA below ops.TC_MIN_ROWS: a hidden unit whose pre-activation lies within fp32 rounding
                # of zero can come out on the other side of the ReLU (tools/acc_diag2.py on this batch: 1 of 203 072
                # elements of layer 3, reference value exactly 0), which moves the gradients below it by one atom's
                # contribution (2.6e-3 of the layer scale here) -- every other quantity agrees to 2e-6
                gate = 6.0 * ref + 2e-5 if not fused_on else max(6.0 * ref + 2e-5, 5e-3)
                assert ours <= gate, (fused_on, k, ours, ref)
        # a step over ~2 500 pairs through the public API (positives by the DataLoader mechanism, vectorised negatives)
        torch.manual_seed(0)
        model = B.Model(data).to(DEV)
        model.train()
        eng = LowerOnlyEngine(data, model)
        sampler = B.RandomSampler(data, 1280)
        ls = [float(eng.train_step(sampler, fast_negatives=True, rng=np.random.default_rng(i))) for i in range(12)]
        assert eng.last_pairs == 2560 and all(np.isfinite(ls)) and np.mean(ls[-3:]) < np.mean(ls[:3])
    finally:
        B.set_flags(B.make_flags(device=DEV))
