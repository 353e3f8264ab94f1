"""ReLU-mask audit of the fused lower stack on the recorded Bi-GNN (GIN+GCN) step: every saved tensor of
bignn_gin_layer_fwd (z, t, y of the five layers over the all-drug merged graph, per-chunk BatchNorm) against an fp64
recomputation, and how many ReLU masks (t > 0, y > 0) differ from the fp64 ones -- a flipped mask moves the gradients
below it by one atom's contribution, which is what the gradient gates of tests/test_gpu_fused_stack.py have to allow.
usage (GPU): python tools/acc_diag3.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bignn_b200 as B
from bignn_b200 import fused
from bignn_b200.engine import BiGNNEngine

DEV = 'cuda:0'


def relu_mask_audit(eng, verbose=True):
    """[(layer, t flips, y flips, max |fp64 pre-activation| at a flip)] of the fused lower stack of `eng` (BiGNNEngine
    with fused_lower=True) for its current parameters; runs one training-mode forward of the stack."""
    mg, spec = eng.merged, eng._stack

    class Ctx:
        needs_input_grad = (False, False, False, True)

    ctx = Ctx()
    params = spec.params()
    bufs = [(l.bn.running_mean.clone(), l.bn.running_var.clone(), l.bn.num_batches_tracked.clone()) for l in spec.layers]
    sink = spec.sink
    if sink is not None:
        n_sink = len(sink)
    fused._GinStack.forward(ctx, spec, True, mg.x, *params)
    torch.cuda.synchronize()
    for l, (m, v, n) in zip(spec.layers, bufs):              # the audit must not advance the running statistics
        l.bn.running_mean.copy_(m); l.bn.running_var.copy_(v); l.bn.num_batches_tracked.copy_(n)
    if sink is not None:
        del sink[n_sink:]

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())

    rp = mg.row_ptr.cpu().numpy()
    ci = mg.col_idx.long()
    rowid = torch.repeat_interleave(torch.arange(mg.A), torch.as_tensor(np.diff(rp))).to(DEV)
    crp = mg.chunk_row_ptr.cpu().numpy()
    seg = torch.repeat_interleave(torch.arange(mg.S), torch.as_tensor(np.diff(crp))).to(DEV)
    cnt = torch.as_tensor(np.diff(crp)).to(DEV).double().unsqueeze(1)
    act = {0: lambda v: v, 1: torch.relu, 2: torch.sigmoid, 3: torch.tanh}
    X = mg.x.double()
    out = []
    for li, (Y, Z, T, mean, rstd, a_in, a_out, din) in enumerate(ctx.saved):
        W1, b1, W2, b2, g, be = [p.detach().double() for p in params[6 * li:6 * li + 6]]
        xin = X
        keep = rowid != ci                                   # remove_self_loops
        agg = torch.zeros(mg.A, xin.shape[1], dtype=torch.float64, device=DEV).index_add_(0, rowid[keep], xin[ci[keep]])
        zref = (1.0 + spec.layers[li]._eps_value()) * xin + agg
        tpre = zref[:, :din] @ W1.t() + b1
        tref = act[a_in](tpre)
        ypre = tref @ W2.t() + b2
        yref = act[a_out](ypre)
        mt = (T > 0) != (tpre > 0) if a_in == 1 else torch.zeros_like(T, dtype=torch.bool)
        my = (Y > 0) != (ypre > 0) if a_out == 1 else torch.zeros_like(Y, dtype=torch.bool)
        nt, ny = int(mt.sum()), int(my.sum())
        worst = max(float(tpre[mt].abs().max()) if nt else 0.0, float(ypre[my].abs().max()) if ny else 0.0)
        out.append((li, nt, ny, worst))
        if verbose:
            print('layer %d  z %.2e  t %.2e  y %.2e | ReLU masks != fp64: t %d  y %d  of %d (max |fp64 pre-activation| '
                  'at a flip %.2e)' % (li, rel(Z[:, :din], zref[:, :din]), rel(T, tref), rel(Y, yref), nt, ny, T.numel(), worst))
        s1 = torch.zeros(mg.S, 64, dtype=torch.float64, device=DEV).index_add_(0, seg, yref)
        mu = s1 / cnt
        var = torch.zeros(mg.S, 64, dtype=torch.float64, device=DEV).index_add_(0, seg, (yref - mu[seg]) ** 2) / cnt
        X = (yref - mu[seg]) / torch.sqrt(var[seg] + spec.layers[li].bn.eps) * g + be
    return out


if __name__ == '__main__':
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
    z = np.load(gold + '/bignn_gin_gcn_step.npz')
    B.set_flags(B.make_flags(device=DEV))
    data = B.BiGNNData.from_npz(gold + '/drugbank_packed.npz', device=DEV)
    model = B.Model(data).to(DEV)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    for k in z.files:
        if k.startswith('sd_init/'):
            sd[k[len('sd_init/'):]] = torch.from_numpy(np.asarray(z[k]))
    model.load_state_dict(sd, strict=False)
    model.train()
    relu_mask_audit(BiGNNEngine(data, model, use_cuda_graph=False, fused_lower=True))
