"""The fused gather-and-score decoder (`bignn_pair_decoder_fwd/bwd`: normalise + gather + concat + MLP + head + loss in
one launch each way) against torch in fp64 and against the layer-by-layer kernels it replaces (`-m gpu`)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200 import ops
from bignn_b200.graph import entry_csr

DEV = 'cuda:0'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def reference(h, ids, Ws, bs, head, target):
    z = F.normalize(h, p=2, dim=1)
    x = torch.cat([z[ids[:, 0].long()], z[ids[:, 1].long()]], 1)
    for l, (w, b) in enumerate(zip(Ws, bs)):
        x = F.linear(x, w, b)
        if l < len(Ws) - 1:
            x = torch.relu(x)
    if head == 0:
        s = torch.sigmoid(x)
        return s, F.binary_cross_entropy(s.view(-1), target.to(s.dtype))
    if head == 1:
        return x, F.binary_cross_entropy_with_logits(x.view(-1), target.to(x.dtype))
    return x, F.cross_entropy(x, target.long())


@pytest.mark.parametrize('P,N,D,widths,head', [(128, 1309, 64, (16, 2, 1), 0), (128, 3242, 64, (16, 3), 2),
                                               (1, 10, 64, (16, 2, 1), 0), (1000, 500, 64, (16, 2, 1), 1),
                                               (5000, 20000, 64, (16, 3), 2), (77, 300, 32, (8, 1), 0),
                                               (333, 300, 128, (16, 4, 1), 0)])
def test_fused_decoder_vs_fp64_torch(P, N, D, widths, head):
    B._lib.load()
    g = torch.Generator().manual_seed(P + N)
    h = torch.randn(N, D, generator=g)
    h[3] = 0.0                                                  # a zero row: the clamp branch of F.normalize
    ids = torch.randint(0, N, (P, 2), generator=g, dtype=torch.int32)
    if P > 2:
        ids[0, 0] = 3
    dims = [2 * D] + list(widths)
    Ws = [torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5 for i in range(len(widths))]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(len(widths))]
    target = torch.randint(0, widths[-1] if head == 2 else 2, (P,), generator=g)
    assert ops.pair_decoder_supported(D, widths)
    # fp64 torch
    h64 = h.double().requires_grad_(True)
    W64 = [w.double().requires_grad_(True) for w in Ws]
    b64 = [b.double().requires_grad_(True) for b in bs]
    s64, l64 = reference(h64, ids, W64, b64, head, target.double() if head != 2 else target)
    l64.backward()
    # fused kernels
    hd = h.to(DEV).requires_grad_(True)
    Wd = [w.to(DEV).requires_grad_(True) for w in Ws]
    bd = [b.to(DEV).requires_grad_(True) for b in bs]
    idd = ids.to(DEV)
    csr = entry_csr(ids.numpy(), N, DEV)
    tgt = target.to(DEV).to(torch.int32 if head == 2 else torch.float32)
    for rep in range(2):                                         # twice: the completion counter resets itself
        for t in [hd] + Wd + bd:
            t.grad = None
        scores, loss = ops.pair_decoder(hd, idd, csr, head, tgt, [t for wb in zip(Wd, bd) for t in wb])
        loss.backward()
        assert rel(scores, s64) < 2e-6
        assert abs(float(loss) - float(l64)) < 2e-6 * max(1.0, abs(float(l64)))
        assert rel(hd.grad, h64.grad) < 5e-6
        for a, b in zip(Wd + bd, W64 + b64):
            assert rel(a.grad, b.grad) < 1e-5          # (fp32 sums over up to 5 000 pairs, per warp then in warp order)
    # scores only (evaluation: no targets)
    s2, _ = ops.pair_decoder(hd.detach(), idd, csr, head, None, [t.detach() for wb in zip(Wd, bd) for t in wb])
    assert torch.equal(s2, scores)


def test_engine_step_with_and_without_fused_decoder(golden_dir, step_golden):
    """the whole train step: fused decoder (default) vs the layer-by-layer decoder kernels vs the golden step."""
    from bignn_b200.engine import BiGNNEngine, _StaticPairBatch
    z = step_golden
    res = {}
    for fused in (True, False):
        if fused:
            os.environ.pop('BIGNN_NO_FUSED_DECODER', None)
        else:
            os.environ['BIGNN_NO_FUSED_DECODER'] = '1'
        try:
            B.set_flags(B.make_flags(device=DEV))
            data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
            model = B.Model(data).to(DEV)
            sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
            for k in z.files:
                if k.startswith('sd_init/'):
                    sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
            model.load_state_dict(sd, strict=False)
            model.train()
            eng = BiGNNEngine(data, model, use_cuda_graph=False)
            st, P = eng.stage_pairs(z['batch_gids'], z['y_true'].astype(np.float32))
            sb = _StaticPairBatch(data, P, data.device)
            sb.load(st)
            model.zero_grad()
            n0 = B._lib.launch_count()
            loss = eng.forward(sb)
            loss.backward()
            res[fused] = dict(loss=float(loss), preds=sb.preds.clone(), launches=B._lib.launch_count() - n0,
                              grads={k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        finally:
            os.environ.pop('BIGNN_NO_FUSED_DECODER', None)
    a, b = res[True], res[False]
    assert abs(a['loss'] - float(z['loss'])) < 1e-5 and abs(a['loss'] - b['loss']) < 1e-6
    assert rel(a['preds'].view(-1), torch.from_numpy(z['pair_preds']).view(-1)) < 1e-5
    assert a['launches'] < b['launches'] - 10
    for k in a['grads']:
        if int(k.split('.')[1]) >= 7:                       # upper level + scorer: well conditioned
            # (on the gradient scale of the layer: a bias in front of a BatchNorm has a true gradient of exactly 0)
            sc = max(float(b['grads'][q].abs().max()) for q in b['grads'] if q.split('.')[1] == k.split('.')[1])
            assert float((a['grads'][k] - b['grads'][k]).abs().max()) <= 2e-5 * max(sc, 1e-12), k
    print('launches per step: fused decoder %d, layer-by-layer decoder %d' % (a['launches'], b['launches']))


def test_gated_readout_equals_gate_mul_then_sum():
    """`bignn_readout_gated_fwd/bwd` (sigmoid(gate) * weight summed per graph in one launch) = gate_mul + sum readout."""
    B._lib.load()
    g = torch.Generator().manual_seed(5)
    sizes = torch.randint(1, 60, (700,), generator=g)
    sizes[10] = 457
    seg = torch.cat([torch.zeros(1, dtype=torch.long), sizes.cumsum(0)]).to(torch.int32).to(DEV)
    A, G = int(sizes.sum()), sizes.numel()
    gate = torch.randn(A, 64, generator=g).to(DEV).requires_grad_(True)
    w = torch.randn(A, 64, generator=g).to(DEV).requires_grad_(True)
    dout = torch.randn(G, 64, generator=g).to(DEV)
    out = ops.gated_readout(gate, w, seg, G)
    out.backward(dout)
    g1, w1 = gate.grad.clone(), w.grad.clone()
    gate.grad = w.grad = None
    ref = ops.readout([ops.gate_mul(gate, w)], seg, G, 'sum')
    ref.backward(dout)
    assert torch.equal(out, ref)
    assert torch.equal(g1, gate.grad) and torch.equal(w1, w.grad)
    want = torch.zeros(G, 64, dtype=torch.float64).index_add_(
        0, torch.repeat_interleave(torch.arange(G), sizes), (torch.sigmoid(gate.detach().double()) * w.detach().double()).cpu())
    assert rel(out, want) < 2e-6
