"""Fused activation-backward paths (`-m gpu`): the derivative of the activation in front of a BatchNorm folded into
the BatchNorm backward (bignn_bn_seg_bwd input_act) and the derivative of the GIN MLP's inner activation folded
into the epilogue of the backward-input tensor-core GEMM (bignn_gemm_tc_masked_f32), against plain torch autograd
of the reference's operator sequence (model/layers.py:27-29,55-57: Linear, act, Linear, act, BatchNorm1d) in fp64,
and against the unfused kernels (identical bits: the same formulas applied to the same values)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200 import ops

DEV = 'cuda:0'


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('M,N,K,nk', [(3712, 64, 64, True), (95038, 64, 64, False), (1000, 64, 48, True),
                                      (129, 56, 36, False), (4097, 32, 64, False), (700, 128, 64, True),
                                      (513, 50, 20, True)])
@pytest.mark.parametrize('act', ['relu', 'sigmoid', 'tanh'])
def test_masked_gemm_equals_gemm_then_act_bwd(M, N, K, nk, act):
    B._lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = torch.randn((N, K) if nk else (K, N), generator=g).to(DEV)
    pre = torch.randn(M, N, generator=g)
    y = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](pre).to(DEV)
    code = ops.ACT_CODES[act]
    plain = ops.gemm_tc(a, b, nk)
    want = ops.act_bwd(y, plain, code)                       # the unfused pair of kernels
    got = ops.gemm_tc(a, b, nk, mask_y=y, mask_act=code)
    assert torch.equal(got, want)
    if act == 'relu':
        ref = (a.double() @ (b.double().t() if nk else b.double())) * (y > 0)
        assert rel(got, ref) < 1e-6 * max(1.0, np.sqrt(K) / 8)


@pytest.mark.parametrize('act', ['relu', 'tanh', 'sigmoid', 'identity'])
@pytest.mark.parametrize('rows,F', [(3040, 64), (95038, 40), (700, 49)])
def test_gin_mlp_bn_block_fused_backward_vs_torch_fp64(act, rows, F):
    """z -> Linear -> act -> Linear -> act -> BatchNorm (3 chunks), fused flags on, vs torch autograd in fp64"""
    B._lib.load()
    g = torch.Generator().manual_seed(rows + F)
    z = torch.randn(rows, F, generator=g)
    w1, b1 = torch.randn(64, F, generator=g) * 0.2, torch.randn(64, generator=g) * 0.1
    w2, b2 = torch.randn(64, 64, generator=g) * 0.2, torch.randn(64, generator=g) * 0.1
    gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    dy = torch.randn(rows, 64, generator=g)
    ptr = np.asarray([0, rows // 3, rows // 2, rows])
    f = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh, 'identity': lambda v: v}[act]
    # ---- reference in fp64
    P = [t.double().requires_grad_(True) for t in (z, w1, b1, w2, b2, gamma, beta)]
    t = f(P[0] @ P[1].t() + P[2])
    x = f(t @ P[3].t() + P[4])
    outs = []
    for i in range(3):
        xs = x[ptr[i]:ptr[i + 1]]
        mu, var = xs.mean(0), xs.var(0, unbiased=False)
        outs.append((xs - mu) / torch.sqrt(var + 1e-5) * P[5] + P[6])
    want = torch.cat(outs)
    want.backward(dy.double())
    # ---- kernels, fused and unfused
    code = ops.ACT_CODES[act]
    seg = torch.as_tensor(ptr.astype(np.int32)).to(DEV)
    res = {}
    for fused in (False, True):
        Q = [t.to(DEV).requires_grad_(True) for t in (z, w1, b1, w2, b2, gamma, beta)]
        inner = code if fused else 0
        tt = ops.linear_act(Q[0], Q[1], Q[2], code, 'oi', bool(inner))
        xx = ops.linear_act(tt, Q[3], Q[4], code, 'oi', fused and code != 0, inner)
        yy = ops.seg_batch_norm(xx, Q[5], Q[6], seg, 3, None, None, None, 1e-5, 0.1, None, code if fused else 0)
        yy.backward(dy.to(DEV))
        res[fused] = (yy.detach(), [q.grad for q in Q])
        assert rel(yy, want) < 5e-6
        scale = {}
        for name, q, p in zip(('dz', 'dw1', 'db1', 'dw2', 'db2', 'dgamma', 'dbeta'), Q, P):
            # a bias that reaches the BatchNorm through identity activations has a true gradient of exactly 0
            # (what is stored is rounding noise): measure biases on the scale of their layer's weight gradient
            sc = float(p.grad.abs().max())
            scale[name] = sc
            if name in ('db1', 'db2'):
                sc = max(sc, scale['dw' + name[-1]])
            err = float((q.grad.double().cpu() - p.grad).abs().max()) / max(sc, 1e-30)
            assert err < 5e-5, (name, fused, err)
    assert torch.equal(res[True][0], res[False][0])
    for a, b in zip(res[True][1], res[False][1]):
        assert torch.equal(a, b)                 # same formulas on the same values: identical bits
