import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, numpy as np
import bignn_b200 as B
from bignn_b200 import ops
B._lib.load()
for (M, Np, Nq, cs) in [(128, 64, 64, 0), (256, 64, 64, 0), (95038, 64, 64, 0), (95038, 64, 40, 1), (3242, 40, 64, 1), (128, 16, 8, 0)]:
    g = torch.Generator().manual_seed(M + Np + Nq)
    p = torch.randn(M, Np, generator=g); q = torch.randn(M, Nq, generator=g) + 0.5
    want = p.double().t() @ q.double()
    d, c = ops.dw_tc(p.cuda(), q.cuda(), cs)
    simt = ops.gemm(p.cuda(), q.cuda(), True, False)
    err = (d.double().cpu() - want)
    print(M, Np, Nq, 'tc max|err|/max|want|', float(err.abs().max() / want.abs().max()), 'simt', float((simt.double().cpu() - want).abs().max() / want.abs().max()),
          'colsum rel', float(((c.double().cpu() - (p if cs == 0 else q).double().sum(0)).abs().max()) / (p if cs == 0 else q).double().sum(0).abs().max()) if c is not None else None)
    if M <= 256:
        print('  first row got ', d[0, :6].cpu().numpy(), '\n  first row want', want[0, :6].numpy())
# timing
p = torch.randn(95038, 64, device='cuda'); q = torch.randn(95038, 64, device='cuda')
for name, fn in (('tc', lambda: ops.dw_tc(p, q, 0)), ('simt', lambda: (ops.gemm(p, q, True, False), ops.colsum(p)))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, 'us per call', e0.elapsed_time(e1) * 100)
