"""Parity tests proper (`-m gpu`): every C-ABI entry point, called through the real CUDA
library on a B200, against the CPU oracle (oracle/bignn_oracle.py, plain torch fp32/fp64) on the
same seeded inputs.  Integer/index outputs must be bit-exact; fp32 outputs are compared with the
tolerance written at each assert."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200 import ops
from oracle import bignn_oracle as O

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def data(golden_dir):
    assert torch.cuda.is_available()
    B._lib.load()
    B.set_flags(B.make_flags(device=DEV))
    return B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)


def random_sym_csr(n, avg_deg, seed, self_loops=False):
    rng = np.random.default_rng(seed)
    m = int(n * avg_deg / 2)
    a = rng.integers(0, n, m)
    b = rng.integers(0, n, m)
    if not self_loops:
        keep = a != b
        a, b = a[keep], b[keep]
    key = np.unique(np.concatenate([a * n + b, b * n + a]))
    row, col = key // n, key % n
    return row, col


# ----------------------------------------------------------------------------- merge
@pytest.mark.parametrize('chunk', [0, 10])
def test_merge_build_bit_exact_vs_reference_chunks(data, drugbank, step_golden, chunk):
    z = step_golden
    gids = z['chunk%d/gids' % chunk]
    rows = [data.gs_map[int(g)] for g in gids]
    m = B.MergedGraph(data.packed, rows)
    assert np.array_equal(m.edge_index.cpu().numpy(), z['chunk%d/edge_index' % chunk].astype(np.int64))
    assert np.array_equal(m.batch.cpu().numpy(), z['chunk%d/batch' % chunk].astype(np.int64))
    assert np.array_equal(m.x.cpu().numpy(), z['chunk%d/x_u8' % chunk].astype(np.float32))
    il = z['chunk%d/ind_list' % chunk]
    assert np.array_equal(m.seg_ptr.cpu().numpy(), np.concatenate([il[:, 0], il[-1:, 1]]))
    el = z['chunk%d/edge_ind_list' % chunk]
    assert np.array_equal(m.edge_ptr.cpu().numpy(), np.concatenate([el[:, 0], el[-1:, 1]]))
    assert np.array_equal(np.diff(m.seg_ptr.cpu().numpy()), z['chunk%d/graph_sizes' % chunk])


@pytest.mark.parametrize('G,seed', [(1, 0), (2, 1), (1309, 2), (1025, 3), (5000, 4), (70000, 5)])
def test_merge_build_bit_exact_vs_oracle(data, drugbank, G, seed):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, drugbank.N, G) if G != 1309 else np.arange(1309)
    m = B.MergedGraph(data.packed, rows)
    o = O.merge_graphs(drugbank, drugbank.gids[rows])
    assert np.array_equal(m.edge_index.cpu().numpy(), o['edge_index'])
    assert np.array_equal(m.batch.cpu().numpy(), o['batch'])
    assert np.array_equal(m.x.cpu().numpy(), o['x'])
    assert np.array_equal(m.seg_ptr.cpu().numpy()[:-1], o['ind_list'][:, 0])
    assert np.array_equal(m.edge_ptr.cpu().numpy()[:-1], o['edge_ind_list'][:, 0])
    # CSR view == sorted COO
    rp = m.row_ptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == m.E
    assert np.array_equal(np.repeat(np.arange(m.A), np.diff(rp)), o['edge_index'][0])
    assert np.array_equal(m.col_idx.cpu().numpy(), o['edge_index'][1])


def test_merge_build_empty(data):
    m = B.MergedGraph(data.packed, [])
    assert m.A == 0 and m.E == 0
    assert m.seg_ptr.cpu().tolist() == [0] and m.row_ptr.cpu().tolist() == [0]


# ----------------------------------------------------------------------------- spmm
def csr_from_coo(row, col, n):
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, row + 1, 1)
    return ops.CSR(torch.as_tensor(np.cumsum(ptr).astype(np.int32)).to(DEV),
                   torch.as_tensor(col.astype(np.int32)).to(DEV), n)


@pytest.mark.parametrize('n,deg,D,self_loops', [(1000, 2.2, 64, False), (777, 2.2, 49, False), (500, 40, 320, False),
                                                (300, 5, 32, True), (200, 3, 16, False), (100, 3, 7, True),
                                                (64, 6, 512, False), (64, 6, 516, False), (50, 4, 130, False)])
def test_spmm_gin_bit_exact(n, deg, D, self_loops):
    row, col = random_sym_csr(n, deg, n + D, self_loops)
    csr = csr_from_coo(row, col, n)
    g = torch.Generator().manual_seed(D)
    x = torch.randn(n, D, generator=g)
    ei = torch.from_numpy(np.stack([row, col]))
    keep = ei[0] != ei[1]
    agg = torch.zeros(n, D).index_add_(0, ei[1][keep], x[ei[0][keep]])
    want = 1.25 * x + agg
    got = ops.spmm(csr, x.to(DEV), ops.SPMM_GIN, 1.25)
    # same neighbour order, unfused adds: identical bits
    assert torch.equal(got.cpu(), want)
    got_sum = ops.spmm(csr, x.to(DEV), ops.SPMM_SUM)
    assert torch.equal(got_sum.cpu(), torch.zeros(n, D).index_add_(0, ei[1], x[ei[0]]))


@pytest.mark.parametrize('D,act', [(64, 'relu'), (64, 'identity'), (49, 'tanh'), (320, 'sigmoid')])
def test_spmm_gcn_vs_oracle(data, drugbank, D, act):
    g = torch.Generator().manual_seed(D)
    n = drugbank.N
    h = torch.randn(n, D, generator=g)
    bias = torch.randn(D, generator=g)
    P = {'l.conv.weight': torch.eye(D), 'l.conv.bias': bias}
    ei = torch.from_numpy(np.stack([drugbank.ddi_row, drugbank.ddi_col]))
    want = O._act(act, O.gcn_conv(h, ei, P, 'l'))
    csr = data.interaction_combo_nxgraph.csr
    got = ops.spmm(csr, h.to(DEV), ops.SPMM_GCN, 0.0, csr.dinv(), bias.to(DEV), ops.act_code(act))
    err = rel(got, want)
    if err >= 2e-6:
        # DESIGN.md "known open item": seen twice in ~20 fresh-box runs ([49-tanh], 5e-5, first run of a new
        # build).  Still a failure -- but say which side is not reproducible.
        got2 = ops.spmm(csr, h.to(DEV), ops.SPMM_GCN, 0.0, csr.dinv(), bias.to(DEV), ops.act_code(act))
        want2 = O._act(act, O.gcn_conv(h, ei, P, 'l'))
        d = (got.cpu() - want).abs()
        pytest.fail('GCN SpMM vs oracle: err {:.3g}; kernel reproducible: {}, oracle reproducible: {}, second '
                    'kernel run vs oracle {:.3g}, kernel vs second oracle run {:.3g}; worst element {} of row degree {}'
                    .format(err, torch.equal(got, got2), torch.equal(want, want2), rel(got2, want), rel(got, want2),
                            np.unravel_index(int(d.argmax()), d.shape),
                            int(np.bincount(drugbank.ddi_row, minlength=n)[int(d.argmax()) // D])))
    # deg^-1/2 itself is bit-exact
    row, col = ei
    deg = torch.zeros(n).index_add_(0, row, torch.ones(row.shape[0])) + 1
    assert torch.equal(csr.dinv().cpu(), deg.pow(-0.5))


@pytest.mark.parametrize('mode', ['gin', 'gcn', 'sum'])
def test_spmm_long_rows_planned(mode):
    """hub rows (degree >> 128) go through the work-item variant; same result as the oracle."""
    n, D = 3000, 64
    rng = np.random.default_rng(5)
    hub = np.concatenate([np.zeros(2500, np.int64), np.ones(700, np.int64), rng.integers(0, n, 20000)])
    oth = np.concatenate([rng.integers(1, n, 2500), rng.integers(2, n, 700), rng.integers(0, n, 20000)])
    keep = hub != oth
    key = np.unique(np.concatenate([hub[keep] * n + oth[keep], oth[keep] * n + hub[keep]]))
    row, col = key // n, key % n
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, row + 1, 1)
    ptr = np.cumsum(ptr)
    csr = ops.CSR(torch.as_tensor(ptr.astype(np.int32)).to(DEV), torch.as_tensor(col.astype(np.int32)).to(DEV), n,
                  row_ptr_host=ptr)
    assert csr.plan is not None and csr.plan.n_multi >= 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, D, generator=g)
    ei = torch.from_numpy(np.stack([row, col]))
    if mode == 'gin':
        want = 1.5 * x.double() + torch.zeros(n, D, dtype=torch.float64).index_add_(0, ei[1], x.double()[ei[0]])
        got = ops.spmm(csr, x.to(DEV), ops.SPMM_GIN, 1.5)
    elif mode == 'sum':
        want = torch.zeros(n, D, dtype=torch.float64).index_add_(0, ei[1], x.double()[ei[0]])
        got = ops.spmm(csr, x.to(DEV), ops.SPMM_SUM)
    else:
        bias = torch.randn(D, generator=g)
        P = {'l.conv.weight': torch.eye(D, dtype=torch.float64), 'l.conv.bias': bias.double()}
        want = torch.relu(O.gcn_conv(x.double(), ei, P, 'l'))
        got = ops.spmm(csr, x.to(DEV), ops.SPMM_GCN, 0.0, csr.dinv(), bias.to(DEV), 1)
    assert rel(got, want) < 2e-6
    assert torch.equal(got, ops.spmm(csr, x.to(DEV), {'gin': 1, 'sum': 0, 'gcn': 2}[mode],
                                     1.5 if mode == 'gin' else 0.0, csr.dinv() if mode == 'gcn' else None,
                                     bias.to(DEV) if mode == 'gcn' else None, 1 if mode == 'gcn' else 0))  # deterministic


def test_spmm_known_answers_small_graphs():
    """P3 path, star + second component + isolated node, self-loop input: results written by hand."""
    # P3: 0 - 1 - 2, x = [1, 10, 100] in every column
    csr = csr_from_coo(np.asarray([0, 1, 1, 2]), np.asarray([1, 0, 2, 1]), 3)
    x = torch.tensor([[1.0], [10.0], [100.0]]).repeat(1, 64).to(DEV)
    assert ops.spmm(csr, x, ops.SPMM_GIN, 1.0)[:, 0].tolist() == [11.0, 111.0, 110.0]
    assert ops.spmm(csr, x, ops.SPMM_SUM)[:, 7].tolist() == [10.0, 101.0, 10.0]
    d = torch.tensor([2.0, 3.0, 2.0]).pow(-0.5)
    want = torch.stack([d[0] * d[0] * 1 + d[0] * d[1] * 10, d[1] * d[0] * 1 + d[1] * d[1] * 10 + d[1] * d[2] * 100,
                        d[2] * d[1] * 10 + d[2] * d[2] * 100])
    got = ops.spmm(csr, x, ops.SPMM_GCN, 0.0, csr.dinv())[:, 63].cpu()
    assert torch.allclose(got, want, rtol=1e-6)
    # star (0; 1,2,3) + edge 4-5 + isolated 6, and an explicit self loop on node 4 that GIN/GCN must drop
    row = np.asarray([0, 0, 0, 1, 2, 3, 4, 4, 5]); col = np.asarray([1, 2, 3, 0, 0, 0, 4, 5, 4])
    csr = csr_from_coo(row, col, 7)
    x = torch.arange(1.0, 8.0).view(7, 1).repeat(1, 16).to(DEV)
    assert ops.spmm(csr, x, ops.SPMM_GIN, 1.0)[:, 3].tolist() == [10.0, 3.0, 4.0, 5.0, 11.0, 11.0, 7.0]
    assert ops.spmm(csr, x, ops.SPMM_SUM)[:, 3].tolist() == [9.0, 1.0, 1.0, 1.0, 11.0, 5.0, 0.0]   # SUM keeps the loop
    got = ops.spmm(csr, x, ops.SPMM_GCN, 0.0, csr.dinv())[:, 0].cpu()
    assert abs(float(got[6]) - 7.0) < 1e-6 and abs(float(got[4]) - (0.5 * 5 + 0.5 * 6)) < 1e-6
    assert torch.equal(csr.dinv().cpu(), torch.tensor([4.0, 2.0, 2.0, 2.0, 2.0, 2.0, 1.0]).pow(-0.5))


def test_spmm_is_its_own_transpose_at_scale():
    """size-independent property on a >L2 operand: <y, A x> == <A y, x> for a symmetric graph,
    and A*ones == degree (+ self coefficient) exactly."""
    n, D = 2_000_000, 64
    row, col = random_sym_csr(n, 2.2, 7)
    csr = csr_from_coo(row, col, n)
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(n, D, device=DEV, generator=g)
    y = torch.randn(n, D, device=DEV, generator=g)
    ax = ops.spmm(csr, x, ops.SPMM_GIN, 1.0)
    ay = ops.spmm(csr, y, ops.SPMM_GIN, 1.0)
    a = (y.double() * ax.double()).sum().item()
    b = (ay.double() * x.double()).sum().item()
    assert abs(a - b) < 1e-6 * max(1.0, abs(a))      # both sides are sums of fp32-rounded rows
    ones = torch.ones(n, D, device=DEV)
    deg = torch.as_tensor(np.bincount(row, minlength=n).astype(np.float32)).to(DEV)
    assert torch.equal(ops.spmm(csr, ones, ops.SPMM_GIN, 1.0), (deg + 1).view(-1, 1).expand(n, D))


# ----------------------------------------------------------------------------- dense
@pytest.mark.parametrize('ta,tb,M,N,K', [(0, 1, 3712, 64, 49), (0, 1, 3712, 64, 64), (0, 0, 1309, 64, 320),
                                         (0, 0, 100, 49, 64), (0, 1, 128, 16, 128), (0, 1, 128, 1, 2),
                                         (1, 0, 64, 49, 36816), (1, 0, 320, 64, 1309), (1, 1, 33, 17, 65),
                                         (0, 1, 1, 1, 1), (1, 0, 64, 64, 700)])
def test_gemm_vs_torch(ta, tb, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((K, M) if ta else (M, K), generator=g)
    b = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    want = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double())
    got = ops.gemm(a.to(DEV), b.to(DEV), bool(ta), bool(tb))
    # fp32 FMA accumulation against an fp64 product: K * 2^-24 worst case, sqrt(K) typical
    assert rel(got, want) < 2e-6 * max(1.0, np.sqrt(K) / 8)
    got2 = ops.gemm(a.to(DEV), b.to(DEV), bool(ta), bool(tb), bias.to(DEV), ops.ACT_CODES['relu'])
    assert rel(got2, torch.relu(want + bias.double())) < 2e-6 * max(1.0, np.sqrt(K) / 8)


@pytest.mark.parametrize('M,N,K,nk', [(3712, 64, 64, True), (95038, 64, 40, True), (1309, 64, 96, False),
                                      (1000, 64, 48, True), (130, 16, 64, True), (4097, 32, 64, False),
                                      (700, 128, 64, True), (128, 64, 8, True), (129, 56, 36, False),
                                      (40000, 64, 64, False), (513, 128, 20, True)])
def test_gemm_tensor_core_3xtf32_vs_fp64(M, N, K, nk):
    """tcgen05 / TMEM path: fp32-level accuracy from three TF32 products."""
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn((N, K) if nk else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    want = a.double() @ (b.double().t() if nk else b.double())
    got = ops.gemm_tc(a.to(DEV), b.to(DEV), nk)
    assert rel(got, want) < 1e-6 * max(1.0, np.sqrt(K) / 8)       # fp32-FMA level (split accumulators)
    got2 = ops.gemm_tc(a.to(DEV), b.to(DEV), nk, bias.to(DEV), ops.ACT_CODES['relu'])
    assert rel(got2, torch.relu(want + bias.double())) < 3e-6 * max(1.0, np.sqrt(K) / 8)
    assert torch.equal(got, ops.gemm_tc(a.to(DEV), b.to(DEV), nk))       # deterministic


@pytest.mark.parametrize('M,Np,Nq,cs', [(95038, 64, 64, 0), (95038, 64, 40, 0), (3242, 64, 64, 1), (3712, 40, 64, 1),
                                        (513, 64, 64, -1), (128, 16, 8, 0), (40000, 64, 64, 0)])
def test_weight_gradient_tensor_core_vs_fp64(M, Np, Nq, cs):
    g = torch.Generator().manual_seed(M + Np + Nq)
    p = torch.randn(M, Np, generator=g)
    q = torch.randn(M, Nq, generator=g) + 0.5
    want = p.double().t() @ q.double()
    d, c = ops.dw_tc(p.to(DEV), q.to(DEV), cs)
    assert rel(d, want) < 3e-6
    if cs >= 0:
        ref = (p if cs == 0 else q).double().sum(0)
        assert rel(c, ref) < 1e-6
    d2, _ = ops.dw_tc(p.to(DEV), q.to(DEV), cs)
    assert torch.equal(d, d2)                       # deterministic


def test_gemm_is_deterministic_with_split_k():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(36816, 64, generator=g).to(DEV)
    b = torch.randn(36816, 49, generator=g).to(DEV)
    r1 = ops.gemm(a, b, True, False)
    r2 = ops.gemm(a, b, True, False)
    assert torch.equal(r1, r2)


def test_colsum_and_act_bwd():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(36816, 64, generator=g)
    assert rel(ops.colsum(x.to(DEV)), x.double().sum(0)) < 1e-6
    x = torch.randn(5, 3, generator=g)
    assert rel(ops.colsum(x.to(DEV)), x.double().sum(0)) < 1e-6
    for name in ('relu', 'sigmoid', 'tanh', 'identity'):
        pre = torch.randn(1000, 64, generator=g, requires_grad=True)
        y = O._act(name, pre)
        dy = torch.randn(1000, 64, generator=g)
        y.backward(dy)
        got = ops.act_bwd(y.detach().to(DEV), dy.to(DEV), ops.act_code(name))
        assert rel(got, pre.grad) < 1e-6


def test_linear_act_autograd():
    g = torch.Generator().manual_seed(5)
    for layout in ('oi', 'io'):
        x = torch.randn(3040, 49, generator=g, requires_grad=True)
        w = torch.randn((64, 49) if layout == 'oi' else (49, 64), generator=g, requires_grad=True)
        b = torch.randn(64, generator=g, requires_grad=True)
        y = torch.relu((x @ w.t() if layout == 'oi' else x @ w) + b)
        dy = torch.randn(3040, 64, generator=g)
        y.backward(dy)
        xd, wd, bd = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
        yd = ops.linear_act(xd, wd, bd, ops.ACT_CODES['relu'], layout)
        yd.backward(dy.to(DEV))
        assert rel(yd, y) < 2e-6
        assert rel(xd.grad, x.grad) < 5e-6
        assert rel(wd.grad, w.grad) < 5e-6
        assert rel(bd.grad, b.grad) < 5e-6


# ----------------------------------------------------------------------------- batch norm
@pytest.mark.parametrize('sizes,C', [([3712, 3040, 817], 64), ([1309], 64), ([5, 2, 900, 33], 49), ([200000], 64)])
def test_seg_batch_norm_vs_torch(sizes, C):
    g = torch.Generator().manual_seed(len(sizes) * 100 + C)
    n = sum(sizes)
    x = (torch.randn(n, C, generator=g) * 2 + 1).relu()
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    dy = torch.randn(n, C, generator=g)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    # reference: torch BatchNorm1d applied chunk by chunk, in order
    # (in fp64: torch's own fp32 CPU kernel is 3.5e-5 off an fp64 run at n = 200 000)
    bn = torch.nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    xr = x.double().requires_grad_(True)
    outs = [bn(xr[ptr[i]:ptr[i + 1]]) for i in range(len(sizes))]
    want = torch.cat(outs)
    want.backward(dy.double())
    rm = torch.zeros(C, device=DEV)
    rv = torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    xd = x.to(DEV).requires_grad_(True)
    gd = gamma.to(DEV).requires_grad_(True)
    bd = beta.to(DEV).requires_grad_(True)
    seg = torch.as_tensor(ptr.astype(np.int32)).to(DEV)
    got = ops.seg_batch_norm(xd, gd, bd, seg, len(sizes), rm, rv, nbt)
    got.backward(dy.to(DEV))
    assert rel(got, want) < 2e-6
    assert rel(rm, bn.running_mean) < 1e-6 and rel(rv, bn.running_var) < 1e-6
    assert int(nbt) == int(bn.num_batches_tracked)
    assert rel(xd.grad, xr.grad) < 2e-5
    assert rel(gd.grad, bn.weight.grad) < 5e-6 and rel(bd.grad, bn.bias.grad) < 5e-6
    # eval mode
    bn.eval()
    assert rel(ops.bn_eval(x.to(DEV), gd.detach(), bd.detach(), rm, rv), bn(x.double())) < 2e-6


@pytest.mark.parametrize('act', ['identity', 'relu'])
def test_seg_batch_norm_backward_cluster_path(act, monkeypatch):
    """many ragged chunks x 64 channels: the chunk-resident backward (one 8-CTA cluster per chunk, both passes while the
    chunk is in L2; bignn_bn_seg_bwd takes it from 24 chunks on) against the two grid-wide passes and against torch."""
    g = torch.Generator().manual_seed(77)
    sizes = [int(v) for v in torch.randint(2, 4200, (61,), generator=g)] + [3, 9, 4097]
    n, C, S = sum(sizes), 64, len(sizes)
    x = (torch.randn(n, C, generator=g) * 2 + 1)
    if act == 'relu':
        x = x.relu()
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    dy = torch.randn(n, C, generator=g)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    bn = torch.nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    xr = x.double().requires_grad_(True)
    want = torch.cat([bn(xr[ptr[i]:ptr[i + 1]]) for i in range(S)])
    want.backward(dy.double())
    want_dx = xr.grad * (x > 0).double() if act == 'relu' else xr.grad      # input_act: derivative of the producer's ReLU
    seg = torch.as_tensor(ptr.astype(np.int32)).to(DEV)
    res = {}
    for mode in ('1', '0'):
        monkeypatch.setenv('BIGNN_BN_CLUSTER', mode)
        xd = x.to(DEV).requires_grad_(True)
        gd = gamma.to(DEV).requires_grad_(True)
        bd = beta.to(DEV).requires_grad_(True)
        n0 = B._lib.launch_count()
        got = ops.seg_batch_norm(xd, gd, bd, seg, S, torch.zeros(C, device=DEV), torch.ones(C, device=DEV),
                                 torch.zeros((), dtype=torch.int64, device=DEV), 1e-5, 0.1, None, ops.ACT_CODES[act])
        n1 = B._lib.launch_count()
        got.backward(dy.to(DEV))
        res[mode] = (xd.grad.clone(), gd.grad.clone(), bd.grad.clone(), B._lib.launch_count() - n1)
        assert rel(xd.grad, want_dx) < 2e-5
        assert rel(gd.grad, bn.weight.grad) < 5e-6 and rel(bd.grad, bn.bias.grad) < 5e-6
    assert res['1'][3] == 2 and res['0'][3] == 4            # cluster path: chunk kernel + parameter gradients
    assert rel(res['1'][0], res['0'][0]) < 1e-6
    assert rel(res['1'][1], res['0'][1]) < 1e-6 and rel(res['1'][2], res['0'][2]) < 1e-6


# ----------------------------------------------------------------------------- readout
@pytest.mark.parametrize('style', ['avg_pool', 'sum'])
@pytest.mark.parametrize('D', [64, 49])
def test_readout_vs_oracle(data, drugbank, style, D):
    rows = np.arange(200)
    m = B.MergedGraph(data.packed, rows)
    g = torch.Generator().manual_seed(D)
    acts = [torch.randn(m.A, D, generator=g, requires_grad=True) for _ in range(3)]
    batch = torch.from_numpy(np.repeat(np.arange(200), np.diff(m.seg_ptr_host)))
    want = O.readout(acts, batch, 200, style)
    dout = torch.randn(200, 3 * D, generator=g)
    want.backward(dout)
    actd = [a.detach().to(DEV).requires_grad_(True) for a in acts]
    got = ops.readout(actd, m.seg_ptr, 200, style)
    got.backward(dout.to(DEV))
    assert rel(got, want) < 1e-6
    for a, b in zip(actd, acts):
        assert rel(a.grad, b.grad) < 1e-6
    # scatter to dataset rows
    perm = torch.randperm(300, generator=g)[:200]
    got2 = ops.readout([a.detach() for a in actd], m.seg_ptr, 200, style,
                       perm.to(torch.int32).to(DEV), 300)
    assert torch.equal(got2[perm.to(DEV)], got.detach())


# ----------------------------------------------------------------------------- decoder
def test_pair_decoder_and_bce_vs_torch():
    g = torch.Generator().manual_seed(11)
    N, D, P = 1309, 64, 128
    h = torch.randn(N, D, generator=g, requires_grad=True)
    ids = torch.randint(0, 120, (P, 2), generator=g)           # heavy row reuse, like a pair batch
    y = (torch.rand(P, generator=g) > 0.5).float()
    w = torch.randn(1, 2 * D, generator=g) * 0.1
    hn = torch.nn.functional.normalize(h, p=2, dim=1)
    z = torch.cat([hn[ids[:, 0]], hn[ids[:, 1]]], 1)
    pred = torch.sigmoid(z @ w.t()).view(-1)
    loss = torch.nn.functional.binary_cross_entropy(pred, y)
    loss.backward()
    hd = h.detach().to(DEV).requires_grad_(True)
    ecsr = B.graph.entry_csr(ids.numpy(), N, DEV)
    zd = ops.pair_gather_norm(hd, ids.to(torch.int32).to(DEV), ecsr)
    pd = ops.linear_act(zd, w.to(DEV), None, ops.ACT_CODES['sigmoid'], 'oi')
    ld = ops.bce(pd.view(-1), y.to(DEV))
    ld.backward()
    assert rel(zd, z) < 1e-6
    assert abs(float(ld) - float(loss)) < 1e-6
    assert rel(hd.grad, h.grad) < 5e-6
    # BCEWithLogits
    x = torch.randn(P, generator=g, requires_grad=True)
    l2 = torch.nn.functional.binary_cross_entropy_with_logits(x, y)
    l2.backward()
    xd = x.detach().to(DEV).requires_grad_(True)
    l2d = ops.bce(xd, y.to(DEV), logits=True)
    l2d.backward()
    assert abs(float(l2d) - float(l2)) < 1e-6 and rel(xd.grad, x.grad) < 1e-6


# ----------------------------------------------------------------------------- errors
def test_argument_errors_raise():
    with pytest.raises(RuntimeError):
        ops.spmm(ops.CSR(torch.zeros(2, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), 1),
                 torch.zeros(1, 4), ops.SPMM_SUM)                      # CPU tensors: no CPU path
    csr = ops.CSR(torch.zeros(2, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV), 1)
    with pytest.raises(RuntimeError):
        B._lib.call('bignn_spmm_f32', csr.row_ptr, csr.col_idx, torch.zeros(1, 4, device=DEV), 4,
                    torch.zeros(1, 4, device=DEV), 4, 1, 4, 9, 0.0, None, None, 0)   # bad mode
    with pytest.raises(ValueError):
        ops.act_code('swish')
