"""Full Bi-GNN (GIN lower + GCN upper) train step on the GPU against the reference-generated
golden vectors (DrugBank fold 1) and the fp64 oracle.

Tolerances: forward quantities 1e-5 relative (max-norm).  Gradients: the reference's own fp32
gradients sit up to 3.5e-4 from an fp64 run of the same code in the lower layers (ill-conditioned
through five BatchNorms; measured in this test), so an fp32 path with another summation order
cannot be within 1e-5 of them (the reference run with 2 instead of 8 CPU threads moves the same
gradients by 4e-3).  The gate is therefore: error against the fp64 oracle no larger than the
reference's own fp32 error against it (+1e-5), per parameter, on the layer's gradient scale; the per-parameter
errors are printed and written to gpurun_out/grad_errors_*.txt (copied to profiles/).
The dense transforms run on the tensor cores as 3xTF32, whose accumulation truncates toward zero
(a ~3e-7 relative bias per transform, profiles/acc_probe.py) where fp32 FMA rounds to nearest."""
GRAD_GATE = 1.0          # round 1: 6.0.  Measured in round 2 (profiles/r2_grad_errors_gin_gcn.txt): this path sits 2.5-10x CLOSER to
                         # the fp64 oracle than the reference's own fp32 values do (worst ratio 0.45), so the gate is the reference's
                         # own error (+1e-5 for parameters whose gradients are at rounding level)
GAT_GRAD_GATE = 2.0      # (see test_gin_gat_step_vs_golden)
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from oracle import bignn_oracle as O

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def world(golden_dir, step_golden):
    B._lib.load()
    B.set_flags(B.make_flags(device=DEV))
    data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
    model = B.Model(data).to(DEV)
    z = step_golden
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    for k in z.files:
        if k.startswith('sd_init/'):
            sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    return data, model


def run_step(data, model, z):
    model.train()
    model.zero_grad()
    B.train._get_initial_embd(data, model)
    bd = B.BatchData(z['positive_gids'], data, sampled_gids=z['sampled_gids'], is_train=False, merge_graphs=False)
    bd.batch_gids = z['batch_gids']
    bd.pair_list = [B.batch.PairRecord(int(l), tuple(g)) for l, g in zip(z['y_true'], z['batch_gids'].tolist())]
    bd.batch_interaction_inds = [data.gs_map[g] for g in bd.batch_gids.flatten().tolist()]
    model.use_layers = 'higher_layers'
    loss = model(bd)
    return bd, loss


def report_gradient_errors(tag, model, g64, z, scale):
    """per-parameter gradient errors on the layer's gradient scale: this path vs the reference's fp32 values, this path
    vs the fp64 oracle, the reference's fp32 values vs the fp64 oracle.  Printed (pytest -s) and written to
    gpurun_out/grad_errors_<tag>.txt so that the distance from north_star's 1e-5 is on record."""
    rows = []
    for k, p in model.named_parameters():
        if not k.startswith('layers.') or p.grad is None:
            continue
        s = scale[k.split('.')[1]]
        g = p.grad.double().cpu().numpy()
        rows.append((k, float(np.abs(g - z['grad/' + k].astype(np.float64)).max()) / s,
                     float(np.abs(g - g64[k]).max()) / s,
                     float(np.abs(z['grad/' + k].astype(np.float64) - g64[k]).max()) / s))
    lines = ['%-44s %12s %12s %12s' % ('parameter (errors / layer gradient scale)', 'vs ref fp32', 'vs fp64', 'ref vs fp64')]
    lines += ['%-44s %12.3e %12.3e %12.3e' % r for r in rows]
    text = '\n'.join(lines)
    print(text)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'grad_errors_%s.txt' % tag), 'w') as f:
            f.write(text + '\n')
    return rows


def test_step_forward_and_gradients(world, step_golden, drugbank, gin_gcn_specs):
    z = step_golden
    data, model = world
    n0 = B._lib.launch_count()
    bd, loss = run_step(data, model, z)
    init_x = data.interaction_combo_nxgraph.init_x
    assert rel(init_x, z['init_x']) < 1e-5
    for l in range(3):
        assert rel(model.acts[l + 2], z['upper/act%d' % (l + 2)]) < 1e-5
    assert rel(model.acts[-2].view(-1), z['pair_preds']) < 1e-5
    assert abs(float(loss) - float(z['loss'])) < 1e-5
    loss.backward()
    assert B._lib.launch_count() - n0 > 100          # the CUDA library did the work
    # fp64 ground truth from the oracle
    sd = O.state_from_npz(z, 'sd0/')
    om = O.OracleModel(gin_gcn_specs, sd, dtype=torch.float64)
    _, _, _, l64 = O.train_step_forward(om, drugbank, z['batch_gids'], z['y_true'])
    l64.backward()
    g64 = {k: v.grad.numpy() for k, v in om.params().items()}
    scale = {}
    for k, g in g64.items():
        lid = k.split('.')[1]
        scale[lid] = max(scale.get(lid, 0.0), float(np.abs(g).max()))
    report_gradient_errors('gin_gcn', model, g64, z, scale)
    worst = 0.0
    for k, p in model.named_parameters():
        if not k.startswith('layers.'):
            continue
        s = scale[k.split('.')[1]]
        ours = float(np.abs(p.grad.double().cpu().numpy() - g64[k]).max()) / s
        ref = float(np.abs(z['grad/' + k].astype(np.float64) - g64[k]).max()) / s
        worst = max(worst, ours)
        assert ours <= GRAD_GATE * ref + 1e-5, (k, ours, ref)
    # upper level + decoder are well conditioned: plain 1e-5 against the reference's fp32
    for k, p in model.named_parameters():
        if k.startswith('layers.') and int(k.split('.')[1]) >= 7:
            s = scale[k.split('.')[1]]
            assert float(np.abs(p.grad.double().cpu().numpy() - z['grad/' + k]).max()) / s < 1e-5, k
    sdm = model.state_dict()
    for k in z.files:
        if k.startswith('sd1/') and 'running' in k:
            assert rel(sdm[k[4:]], z[k]) < 1e-5, k
        if k.startswith('sd1/') and 'num_batches' in k:
            assert int(sdm[k[4:]]) == int(z[k])


def test_lower_chunk_activations(world, step_golden):
    """per-layer activations of the last (29-graph) chunk and pooled rows of the first."""
    z = step_golden
    data, model = world
    model.train()
    model.use_layers = 'lower_layers'
    for c in (0, 10):
        gids = z['chunk%d/batch_gids' % c]
        bd = B.BatchData(gids, data, is_train=False, ignore_pairs=True)
        data.interaction_combo_nxgraph.init_x = torch.zeros((data.N, 320), device=DEV)
        out = model(bd)
        if c == 10:
            for l in range(5):
                assert rel(model.acts[l + 1], z['chunk10/act%d' % (l + 1)]) < 1e-5
            assert rel(out, z['chunk10/act6']) < 1e-5
        else:
            assert rel(out, z['chunk0/pooled']) < 1e-5


def test_gin_gat_step_vs_golden(golden_dir, drugbank):
    """GIN lower + GAT upper (the reference's shipped upper-level type), same gates as above."""
    z = np.load(os.path.join(golden_dir, 'bignn_gin_gat_step.npz'))
    B.set_flags(B.make_flags(device=DEV, higher_level_gnn_type='gat'))
    with open(os.path.join(golden_dir, 'bignn_gin_gat_layers.txt')) as f:
        lines = f.read().split()
    f_ = B.get_flags()
    assert [getattr(f_, 'layer_%d' % i) for i in range(1, f_.layer_num + 1)] == lines
    data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
    model = B.Model(data).to(DEV)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    bd, loss = run_step(data, model, z)
    assert rel(data.interaction_combo_nxgraph.init_x, z['init_x']) < 1e-5
    errs = [rel(model.acts[l + 2], z['upper/act%d' % (l + 2)]) for l in range(3)]
    print('GIN+GAT golden: upper activations vs the reference', ['%.2e' % e for e in errs])
    for e in errs:
        # round 1 needed 2e-5 here (1.06e-5 measured with the 1 309-row upper transforms as 3xTF32); they now run
        # as fp32 FMA (ops.TC_MIN_ROWS) and the stack has to meet north_star's 1e-5
        assert e < 1e-5
    assert abs(float(loss) - float(z['loss'])) < 1e-5
    loss.backward()
    specs = O.parse_specs(lines)
    om = O.OracleModel(specs, O.state_from_npz(z, 'sd0/'), dtype=torch.float64)
    _, _, _, l64 = O.train_step_forward(om, drugbank, z['batch_gids'], z['y_true'])
    l64.backward()
    g64 = {k: v.grad.numpy() for k, v in om.params().items()}
    scale = {}
    for k, g in g64.items():
        scale[k.split('.')[1]] = max(scale.get(k.split('.')[1], 0.0), float(np.abs(g).max()))
    report_gradient_errors('gin_gat', model, g64, z, scale)
    for k, p in model.named_parameters():
        if k.startswith('layers.'):
            s = scale[k.split('.')[1]]
            ours = float(np.abs(p.grad.double().cpu().numpy() - g64[k]).max()) / s
            ref = float(np.abs(z['grad/' + k].astype(np.float64) - g64[k]).max()) / s
            # the GIN+GAT step: worst ratio measured on B200 1.40 (layers.3.conv.nn.2.weight: 5.9e-5 against the
            # reference's own 4.2e-5, profiles/r2_grad_errors_gin_gat.txt), every other parameter below 1.0
            assert ours <= GAT_GRAD_GATE * ref + 1e-5, (k, ours, ref)
    B.set_flags(B.make_flags(device=DEV))
