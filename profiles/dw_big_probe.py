"""Times bignn_dw_tc_f32 (weight + bias gradient of one 64 x 64 Linear) at the C4 size (default 6 M rows, > L2);
the ncu --set full capture of r2 runs this script."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bignn_b200 as B
from bignn_b200 import ops
B._lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
p = torch.randn(M, 64, device='cuda'); q = torch.randn(M, 64, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
for _ in range(3): ops.dw_tc(p, q, 0)
ts = []
for _ in range(8):
    flush.fill_(0.)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.dw_tc(p, q, 0); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.mean(ts))
print(json.dumps({'rows': M, 'ms': round(ms, 4), 'algorithmic_GBps': round(2 * 4.0 * 64 * M / ms / 1e6, 1)}))
