"""Probe of the kernels changed late in round 1, at a > L2 shape (2 M rows x 64): weight-gradient kernel with
64-row tiles / two CTAs per SM, 128-bit BatchNorm passes with the fused activation backward, masked tcgen05 GEMM.
Prints CUDA-event times and achieved algorithmic GB/s; run under `ncu --set full -k regex:...` for the captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import json
import numpy as np
import torch
import bignn_b200 as B
from bignn_b200 import ops
B._lib.load()
M, C = 2_000_000, 64
dev = 'cuda:0'
torch.manual_seed(0)
x = torch.randn(M, C, device=dev).relu_()
g = torch.randn(M, C, device=dev)
w = torch.randn(C, C, device=dev)
gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
S = M // 3840
seg = torch.as_tensor(np.concatenate([np.arange(S) * 3840, [M]]).astype(np.int32)).to(dev)
xr = x.clone().requires_grad_(True)
y = ops.seg_batch_norm(xr, gamma, beta, seg, S, None, None, None, 1e-5, 0.1, None, 1)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


unit = 4.0 * M * C
res = {}
for name, fn, passes in (
        ('dw_tc 64x64 (+colsum)', lambda: ops.dw_tc(g, x, 0), 2),
        ('gemm_tc', lambda: ops.gemm_tc(g, w, True), 2),
        ('gemm_tc masked relu', lambda: ops.gemm_tc(g, w, True, mask_y=x, mask_act=1), 3),
        ('bn fwd (stats + apply)', lambda: ops.seg_batch_norm(x, gamma, beta, seg, S, None, None, None), 3),
        ('bn bwd fused relu (sums + apply)', lambda: torch.autograd.grad(y, xr, g, retain_graph=True), 5),
        ('act_bwd (the pass the fusion removes)', lambda: ops.act_bwd(x, g, 1), 3)):
    ms = timed(fn)
    res[name] = dict(ms=round(ms, 4), algorithmic_GBps=round(passes * unit / ms / 1e6, 1))
    print(name, res[name])
print(json.dumps(res))
