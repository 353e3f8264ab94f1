"""Row-partitioned upper level on the device (`-m gpu`, one GPU): the kernels a rank of an N-GPU job runs
(`bignn_spmm_rows_f32`, `bignn_spmm_planned_rows_f32`, `bignn_bn_rows_*`) are driven here for EVERY
simulated rank of a world of 1-8 on one device -- the rank sum that NCCL performs in the job is done with
a plain fp64 add -- and compared with the unpartitioned kernels (bit-exact for the SpMM: the position map
is monotone, so each row sums its neighbours in the same order) and with the CPU oracle.  The multi-process
plumbing itself is covered by tests/test_dist_gloo.py (gloo, CPU) and the 2-GPU NCCL bench run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200 import ops, _lib
from bignn_b200.graph import InteractionGraph, PartitionedInteractionGraph
from bignn_b200.engine import BiGNNEngine
from bignn_b200 import synthetic as S

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def skewed_graph(n, m, seed):
    _, row, col = S.interaction_graph(n, m, seed)
    return InteractionGraph(n, row, col, DEV)


def to_positions(pg, x):
    """[N, D] in drug-row order -> [n_pad, D] in the gathered position space (what the all-gather builds)."""
    out = torch.zeros((pg.n_pad, x.shape[1]), dtype=x.dtype, device=x.device)
    out[torch.as_tensor(pg.part.pos(np.arange(pg.n))).to(x.device)] = x
    return out


@pytest.mark.parametrize('world', [1, 2, 3, 8])
@pytest.mark.parametrize('mode', [ops.SPMM_GCN, ops.SPMM_GIN, ops.SPMM_SUM])
def test_spmm_rows_equals_unpartitioned_bit_exact(world, mode):
    B._lib.load()
    n, D = 6000, 64
    full = skewed_graph(n, 90_000, 5)             # hub rows > 32 neighbours -> the planned (work-item) kernels
    assert full.csr.plan is not None and full.csr.plan.n_multi > 0
    torch.manual_seed(0)
    x = torch.randn(n, D, device=DEV)
    bias = torch.randn(D, device=DEV)
    dinv = full.csr.dinv() if mode == ops.SPMM_GCN else None
    want = ops.spmm(full.csr, x, mode, 1.25, dinv, bias, 1)
    got = torch.empty_like(want)
    for r in range(world):
        pg = PartitionedInteractionGraph(full, r, world)
        y = ops.spmm(pg.csr, to_positions(pg, x), mode, 1.25, pg.csr.dinv() if mode == ops.SPMM_GCN else None, bias, 1)
        assert y.shape[0] == pg.n_loc
        got[pg.lo:pg.hi] = y
    assert torch.equal(got, want)


@pytest.mark.parametrize('world', [2, 4])
def test_spmm_rows_short_rows_and_odd_width(world):
    """no hub rows (plain sub-warp-per-row kernel) and a width that takes the scalar path"""
    B._lib.load()
    rng = np.random.default_rng(1)
    n = 3000
    a, b = rng.integers(0, n, 9000), rng.integers(0, n, 9000)
    keep = a != b
    key = np.unique(np.concatenate([a[keep] * n + b[keep], b[keep] * n + a[keep]]))
    full = InteractionGraph(n, key // n, key % n, DEV)
    assert full.csr.plan is None
    for D in (64, 50, 7):
        x = torch.randn(n, D, device=DEV)
        want = ops.spmm(full.csr, x, ops.SPMM_GCN, 0.0, full.csr.dinv(), None, 0)
        for r in range(world):
            pg = PartitionedInteractionGraph(full, r, world)
            assert pg.csr.plan is None
            y = ops.spmm(pg.csr, to_positions(pg, x), ops.SPMM_GCN, 0.0, pg.csr.dinv(), None, 0)
            assert torch.equal(y, want[pg.lo:pg.hi]), (D, r)


@pytest.mark.parametrize('world', [1, 2, 5])
def test_bn_rows_equals_segmented_bn_and_oracle(world):
    B._lib.load()
    torch.manual_seed(1)
    n, C = 5003, 64
    x = (torch.randn(n, C, device=DEV) * 3 + 1.5).requires_grad_(True)
    gamma = torch.rand(C, device=DEV) + 0.5
    beta = torch.randn(C, device=DEV)
    dy = torch.randn(n, C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    seg = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    g1, b1 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = ops.seg_batch_norm(x, g1, b1, seg, 1, rm, rv, nbt)
    y_ref.backward(dy)
    # ---- the same batch split over `world` simulated ranks
    bounds = np.linspace(0, n, world + 1).astype(int)
    bounds[1:-1] += 7                       # unequal blocks
    parts = 4
    wsb = _lib.call('bignn_bn_rows_workspace_bytes', C, parts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    xd = x.detach()
    tot = torch.zeros((2, C), dtype=torch.float64, device=DEV)
    for r in range(world):
        xs = xd[bounds[r]:bounds[r + 1]]
        s = torch.empty((2, C), dtype=torch.float64, device=DEV)
        _lib.call('bignn_bn_rows_sums', xs, xs.stride(0), None, 0, xs.shape[0], C, parts, None, None, s, ws, int(wsb))
        tot += s
    y = torch.empty_like(xd)
    means, rstds = [], []
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt2 = torch.zeros((), dtype=torch.int64, device=DEV)
    for r in range(world):
        xs, ys = xd[bounds[r]:bounds[r + 1]], y[bounds[r]:bounds[r + 1]]
        mean, rstd = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        first = r == 0                       # running buffers are replicated: every rank applies the same update
        _lib.call('bignn_bn_rows_fwd_apply', xs, xs.stride(0), ys, ys.stride(0), xs.shape[0], C, parts, tot, n,
                  gamma, beta, 1e-5, 0.1, rm2 if first else None, rv2 if first else None, nbt2 if first else None,
                  mean, rstd)
        means.append(mean); rstds.append(rstd)
    assert rel(y, y_ref) < 1e-6
    assert rel(rm2, rm) < 1e-6 and rel(rv2, rv) < 1e-6 and int(nbt2) == int(nbt) == 1
    totb = torch.zeros((2, C), dtype=torch.float64, device=DEV)
    for r in range(world):
        xs, ds = xd[bounds[r]:bounds[r + 1]], dy[bounds[r]:bounds[r + 1]]
        s = torch.empty((2, C), dtype=torch.float64, device=DEV)
        _lib.call('bignn_bn_rows_sums', xs, xs.stride(0), ds, ds.stride(0), xs.shape[0], C, parts, means[r], rstds[r],
                  s, ws, int(wsb))
        totb += s
    dx = torch.empty_like(xd)
    for r in range(world):
        sl = slice(bounds[r], bounds[r + 1])
        xs, ds, dxs = xd[sl], dy[sl], dx[sl]
        _lib.call('bignn_bn_rows_bwd_apply', xs, xs.stride(0), ds, ds.stride(0), dxs, dxs.stride(0), xs.shape[0], C,
                  parts, gamma, means[r], rstds[r], totb, n, 0)
    assert rel(dx, x.grad) < 2e-6
    assert rel(totb[0].float(), b1.grad) < 1e-6 and rel(totb[1].float(), g1.grad) < 1e-6
    # ---- and against torch's own BatchNorm1d on the CPU (the reference's operator, model/layers.py:57)
    bn = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.copy_(gamma.cpu()); bn.bias.copy_(beta.cpu())
    xc = xd.cpu().clone().requires_grad_(True)
    yc = bn(xc)
    yc.backward(dy.cpu())
    assert rel(y, yc) < 1e-5 and rel(dx, xc.grad) < 1e-5
    assert rel(rm2, bn.running_mean) < 1e-6 and rel(rv2, bn.running_var) < 1e-6


def test_bn_rows_empty_rank_and_errors():
    B._lib.load()
    C, parts = 64, 2
    wsb = _lib.call('bignn_bn_rows_workspace_bytes', C, parts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    s = torch.full((2, C), 7.0, dtype=torch.float64, device=DEV)
    _lib.call('bignn_bn_rows_sums', None, 64, None, 0, 0, C, parts, None, None, s, ws, int(wsb))     # a rank with no rows
    assert float(s.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        _lib.call('bignn_bn_rows_sums', None, 64, None, 0, 0, C, parts, None, None, s, ws, 8)         # workspace too small


def test_engine_partition_path_world1_matches_golden(golden_dir, step_golden):
    """the row-partitioned code path with one rank (row offset 0, no collective) against the reference's
    golden step and the plain engine: loss, predictions and every gradient"""
    from tests.test_gpu_engine import fresh
    from bignn_b200.engine import _StaticPairBatch
    z = step_golden
    res = {}
    for part in (False, True):
        data, model = fresh(golden_dir, z)
        eng = BiGNNEngine(data, model, use_cuda_graph=False, partition_upper=part)
        st, P = eng.stage_pairs(z['batch_gids'], z['y_true'].astype(np.float32))
        sb = _StaticPairBatch(data, P, data.device, eng.upper)
        sb.load(st)
        loss = eng.forward(sb)
        loss.backward()
        assert abs(float(loss) - float(z['loss'])) < 1e-5
        assert rel(sb.preds.view(-1), z['pair_preds']) < 1e-5
        res[part] = (float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    assert abs(res[True][0] - res[False][0]) < 1e-7
    for k, g in res[False][1].items():
        sc = float(g.abs().max())
        if k.endswith('conv.bias'):          # true gradient 0 in front of a BatchNorm: rounding noise
            sc = max(sc, float(res[False][1][k.replace('conv.bias', 'conv.weight')].abs().max()))
        assert float((res[True][1][k] - g).abs().max()) <= 1e-5 * max(sc, 1e-12), k
