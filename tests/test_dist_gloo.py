"""Multi-rank host logic on CPU (world_size 2, gloo): chunk sharding by drug, the pooled-row
exchange and its backward, the flat gradient all-reduce and the replay of BatchNorm running-buffer
updates in global chunk order.  Device arithmetic is the torch-CPU stand-in of tests/fake_backend.py
(the real kernels are exercised by the `gpu` tests); the check is that a 2-rank step equals the
1-rank step and the reference's golden vectors."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, partition_upper=False):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import bignn_b200 as B
    from bignn_b200.engine import BiGNNEngine
    from tests import fake_backend
    fake_backend.install()
    B.set_flags(B.make_flags(device='cpu'))
    gold = os.path.join(ROOT, 'tests', 'golden')
    z = np.load(os.path.join(gold, 'bignn_gin_gcn_step.npz'))
    data = B.BiGNNData.from_npz(os.path.join(gold, 'drugbank_packed.npz'), device='cpu')
    model = B.Model(data)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    for k in z.files:
        if k.startswith('sd_init/'):
            sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
    model.load_state_dict(sd, strict=False)
    model.train()
    eng = BiGNNEngine(data, model, use_cuda_graph=False, rank=rank, world=world, partition_upper=partition_upper)
    # run forward/backward by hand so that gradients can be inspected before Adam
    st, P = eng.stage_pairs(z['batch_gids'], z['y_true'].astype(np.float32))
    from bignn_b200.engine import _StaticPairBatch
    sb = _StaticPairBatch(data, P, data.device, eng.upper)
    sb.load(st)
    loss = eng.forward(sb)
    loss.backward()
    if world > 1:
        eng._sync_lower(sb)
        w, (a, b) = world, eng._chk_sum.tolist()
        assert abs(w * b - a * a) < 0.5, 'ranks disagree on the pair batch'
    ig = eng._ig()
    res = {'loss': float(loss.detach()), 'chunks': np.asarray(eng.my_chunks),
           'init_x': (ig.init_x_full if eng.upper is not None else ig.init_x).detach().numpy()}
    if eng.upper is not None:
        res['upper_rows'] = np.asarray([eng.upper.lo, eng.upper.hi])
        res['preds'] = sb.preds.numpy()
    for k, p in model.named_parameters():
        if k.startswith('layers.'):
            res['grad/' + k] = p.grad.numpy()
    for k, v in model.state_dict().items():
        if k.startswith('layers.') and ('running' in k or 'num_batches' in k):
            res['buf/' + k] = v.numpy()
    np.savez(os.path.join(out_dir, 'w%d_r%d_p%d.npz' % (world, rank, int(partition_upper))), **res)
    if world > 1:
        dist.destroy_process_group()


def test_shard_chunks_covers_and_balances():
    sys.path.insert(0, ROOT)
    from bignn_b200.dist import shard_chunks
    w = [3712, 3040, 3500, 3300, 3600, 3400, 3200, 3100, 3800, 3000, 817]
    for world in (1, 2, 3, 4, 8, 16):
        sh = shard_chunks(w, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        if world <= 4:
            loads = [sum(w[a:b]) for a, b in sh]
            assert max(loads) <= 1.35 * sum(w) / world


def _compare(tmp_path, part, world=2):
    rs = [np.load(os.path.join(tmp_path, 'w%d_r%d_p%d.npz' % (world, r, part))) for r in range(world)]
    r0 = rs[0]
    s = np.load(os.path.join(tmp_path, 'w1_r0_p0.npz'))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'bignn_gin_gcn_step.npz'))
    # the ranks own disjoint, covering chunk ranges
    assert r0['chunks'][0] == 0 and rs[-1]['chunks'][1] == 11
    assert all(a['chunks'][1] == b['chunks'][0] for a, b in zip(rs, rs[1:]))
    for r in rs:
        assert abs(float(r['loss']) - float(s['loss'])) < 1e-6
    assert abs(float(r0['loss']) - float(z['loss'])) < 1e-5
    assert np.abs(r0['init_x'] - s['init_x']).max() < 1e-6
    for k in s.files:
        if k.startswith('grad/'):
            # a bias in front of a BatchNorm has a true gradient of exactly 0: what is stored is rounding noise
            # -> measure it on the gradient scale of the layer's weight
            sc = max(np.abs(s[k]).max(), 1e-6)
            if part and k.endswith('conv.bias'):
                sc = max(sc, np.abs(s[k.replace('conv.bias', 'conv.weight')]).max())
            lid = int(k.split('.')[1])
            tol = 2e-4 if lid < 5 else (2e-5 if part else 1e-6)   # lower grads are re-associated sums over chunks
            for r in rs:
                assert np.abs(r[k] - s[k]).max() / sc < tol, k
            if not part or lid < 5 or '.conv.' in k:
                for r in rs[1:]:
                    assert np.array_equal(r0[k], r[k]), k     # every rank holds the same reduced gradient
            else:
                for r in rs[1:]:
                    assert np.abs(r0[k] - r[k]).max() / sc < 1e-6, k
        if k.startswith('buf/'):
            if 'num_batches' in k:
                assert all(int(r[k]) == int(s[k]) for r in rs), k
            else:
                for r in rs:
                    assert np.abs(r[k] - s[k]).max() < 1e-6, k
    return rs, s


def test_two_rank_step_equals_single_rank_and_golden(tmp_path):
    """drug-sharded lower level, replicated upper level"""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), False), nprocs=2, join=True)
    mp.spawn(_worker, args=(1, port, str(tmp_path), False), nprocs=1, join=True)   # own process: keeps this one's threads
    _compare(tmp_path, 0)


@pytest.mark.parametrize('world', [2, 3])
def test_row_partitioned_upper_level_equals_single_rank_and_golden(tmp_path, world):
    """drug-sharded lower level + interaction-graph rows partitioned by source drug: per-layer all-gather,
    BatchNorm statistics all-reduce, replicated scorer."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), True), nprocs=world, join=True)
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path), False), nprocs=1, join=True)
    rs, s = _compare(tmp_path, 1, world)
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'bignn_gin_gcn_step.npz'))
    assert rs[0]['upper_rows'][0] == 0 and rs[-1]['upper_rows'][1] == 1309
    assert all(a['upper_rows'][1] == b['upper_rows'][0] for a, b in zip(rs, rs[1:]))
    for r in rs:
        assert np.abs(r['preds'].reshape(-1) - z['pair_preds'].reshape(-1)).max() < 1e-5


def test_single_rank_partition_is_the_unpartitioned_path(tmp_path):
    """world = 1 through the row-partitioned code path (row_offset 0, no collective) = the plain path."""
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path), True), nprocs=1, join=True)
    mp.spawn(_worker, args=(1, _free_port(), str(tmp_path), False), nprocs=1, join=True)
    a = np.load(os.path.join(tmp_path, 'w1_r0_p1.npz'))
    b = np.load(os.path.join(tmp_path, 'w1_r0_p0.npz'))
    assert abs(float(a['loss']) - float(b['loss'])) < 1e-7
    for k in b.files:
        if k.startswith('grad/'):
            sc = max(np.abs(b[k]).max(), 1e-6)
            assert np.abs(a[k] - b[k]).max() / sc < 1e-5, k


def test_row_partition_positions():
    sys.path.insert(0, ROOT)
    from bignn_b200.graph import RowPartition
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 50, 1000)
    deg[3] = 5000                                   # a hub
    ptr = np.concatenate([[0], np.cumsum(deg)])
    for world in (1, 2, 4, 8):
        p = RowPartition(ptr, world)
        assert p.bounds[0] == 0 and p.bounds[-1] == 1000 and np.all(np.diff(p.bounds) >= 0)
        pos = p.pos(np.arange(1000))
        assert np.all(np.diff(pos) > 0) and pos.max() < p.n_pad          # monotone, inside the padded space
        for r in range(world):
            lo, hi = p.rows_of(r)
            assert np.array_equal(pos[lo:hi], r * p.n_max + np.arange(hi - lo))
        if world <= 4:
            loads = [ptr[p.bounds[r + 1]] - ptr[p.bounds[r]] for r in range(world)]
            assert max(loads) <= 5000 + 1.3 * ptr[-1] / world


def _traj_worker(rank, world, port, out_dir, arch):
    """several consecutive train steps (Adam included) on the reference's recorded batches"""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    import bignn_b200 as B
    from bignn_b200.engine import BiGNNEngine
    from tests import fake_backend
    fake_backend.install()
    gold = os.path.join(ROOT, 'tests', 'golden')
    if arch == 'drugcombo':
        B.set_flags(B.make_flags(dataset='drugcombo', higher_level_gnn_type='gat', device='cpu'))
        z = np.load(os.path.join(gold, 'bignn_drugcombo_step.npz'))
        s = np.load(os.path.join(gold, 'bignn_drugcombo_sampler_seq.npz'))
        data = B.BiGNNData.from_npz(os.path.join(gold, 'drugcombo_packed.npz'), device='cpu')
    else:
        B.set_flags(B.make_flags(device='cpu'))
        z = np.load(os.path.join(gold, 'bignn_gin_gcn_step.npz'))
        s = np.load(os.path.join(gold, 'bignn_gin_gcn_sampler_seq.npz'))
        data = B.BiGNNData.from_npz(os.path.join(gold, 'drugbank_packed.npz'), device='cpu')
    model = B.Model(data)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    for k in z.files:
        if k.startswith('sd_init/'):
            sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
    model.load_state_dict(sd, strict=False)
    model.train()
    eng = BiGNNEngine(data, model, use_cuda_graph=False, rank=rank, world=world)
    n = int(os.environ.get('BIGNN_TRAJ_STEPS', 5))
    losses = []
    st, P = eng.stage_pairs(z['batch_gids'], z['y_true'].astype(np.float32))
    losses.append(eng.read_loss(eng.step_staged(st, P)))
    for i in range(n):
        gids = np.concatenate([s['pos'][i], s['neg'][i]])
        st, P = eng.stage_pairs(gids, s['y'][i].astype(np.float32))
        losses.append(eng.read_loss(eng.step_staged(st, P)))
    np.savez(os.path.join(out_dir, 'traj_%s_w%d_r%d.npz' % (arch, world, rank)), losses=np.asarray(losses),
             want=np.concatenate([[float(z['loss'])], s['losses'][:n]]))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize('arch', ['drugcombo', 'gin_gcn'])
def test_two_rank_loss_trajectory_follows_the_reference(tmp_path, arch):
    """DrugCombo architecture (MetaLayer upper level: replicated; lower level sharded by drug) and GIN+GCN (upper
    level row-partitioned) over consecutive train steps with Adam on 2 ranks: every rank's loss trajectory = the
    1-rank trajectory = the losses the reference itself recorded for these batches."""
    mp.spawn(_traj_worker, args=(2, _free_port(), str(tmp_path), arch), nprocs=2, join=True)
    mp.spawn(_traj_worker, args=(1, _free_port(), str(tmp_path), arch), nprocs=1, join=True)
    one = np.load(os.path.join(tmp_path, 'traj_%s_w1_r0.npz' % arch))
    two = [np.load(os.path.join(tmp_path, 'traj_%s_w2_r%d.npz' % (arch, r))) for r in range(2)]
    # (the CPU stand-in's threaded torch kernels are not bit-reproducible between processes; the CUDA kernels are)
    assert np.abs(two[0]['losses'] - two[1]['losses']).max() < 2e-6
    dev_ref = np.abs(one['losses'] - one['want'])
    dev_two = np.abs(two[0]['losses'] - one['losses'])
    print(arch, 'loss trajectory: 1 rank vs reference', dev_ref, ' 2 ranks vs 1 rank', dev_two)
    # DrugCombo (GAT upper level) follows the reference to ~1e-6 over all steps.  GIN+GCN is chaotic from step 1 on:
    # its lower-level gradients are ill-conditioned (the reference's own fp32 gradients differ from an fp64 run of the
    # same code by 3.5e-4) and Adam's first updates are lr * g / |g| -- sign-like -- so every
    # re-association of a sum moves some weights by 2 lr; the deviation grows ~4x per step in ANY implementation
    # (measured: 0, 1.2e-5, 3.4e-5, 1.3e-4, 4.2e-4, 1.6e-3 for this path on one rank)
    k = np.arange(len(dev_ref))
    bound = (2e-6 if arch == 'drugcombo' else 2e-5) * 4.0 ** k      # (DrugCombo: 1e-7 .. 4e-5 over six steps)
    assert dev_ref[0] < 1e-5 and np.all(dev_ref <= bound), dev_ref
    assert dev_two[0] < 1e-6 and np.all(dev_two <= bound), dev_two
