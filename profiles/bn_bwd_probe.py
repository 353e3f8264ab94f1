"""Times bignn_bn_seg_bwd at the C4 size (6 M rows x 64, 1 563 chunks): chunk-resident cluster kernel against the two
grid-wide passes (BIGNN_BN_CLUSTER=0)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bignn_b200 as B
from bignn_b200 import ops, _lib
B._lib.load()
A = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
S = A // 3840
ptr = np.linspace(0, A, S + 1).astype(np.int32)
seg = torch.as_tensor(ptr).cuda()
Y = torch.randn(A, 64, device='cuda').relu(); g = torch.randn(A, 64, device='cuda'); dY = torch.empty_like(Y)
gamma = torch.rand(64, device='cuda') + .5
mean = torch.randn(S, 64, device='cuda') * .1 + .4; rstd = torch.rand(S, 64, device='cuda') + .5
dgamma = torch.empty(64, device='cuda'); dbeta = torch.empty(64, device='cuda')
parts = ops.bn_parts(S, A)
wsb = _lib.call('bignn_bn_workspace_bytes', S, 64, parts); ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
res = {'rows': A, 'chunks': S, 'parts': parts}
outs = {}
for mode in ('1', '0'):
    os.environ['BIGNN_BN_CLUSTER'] = mode
    fn = lambda: _lib.call('bignn_bn_seg_bwd', Y, 64, g, 64, dY, 64, seg, S, 64, parts, gamma, mean, rstd, dgamma, dbeta, 1, ws, int(wsb))
    for _ in range(3): fn()
    ts = []
    for _ in range(8):
        flush.fill_(0.)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts))
    outs[mode] = (dY.clone(), dgamma.clone())
    res['cluster' if mode == '1' else 'two passes'] = dict(ms=round(ms, 4), algorithmic_GBps=round(3 * 4.0 * 64 * A / ms / 1e6, 1))
res['max |dY cluster - dY two-pass|'] = float((outs['1'][0] - outs['0'][0]).abs().max())
print(json.dumps(res))
