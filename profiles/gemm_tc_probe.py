"""Times the tcgen05 3xTF32 GEMM (and the SIMT fp32 GEMM) on the path's tall-skinny shape."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bignn_b200 as B
from bignn_b200 import ops
B._lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 95038
a = torch.randn(M, 64, device='cuda'); w = torch.randn(64, 64, device='cuda'); b = torch.randn(64, device='cuda')
out = torch.empty(M, 64, device='cuda')
flush = torch.empty(64 * 1024 * 1024, device='cuda')
res = {}
for name, fn in (('tc', lambda: ops.gemm_tc(a, w, True, b, 1, out=out)), ('simt', lambda: ops.gemm(a, w, False, True, b, 1, out=out))):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    res[name] = dict(us=sorted(ts)[len(ts) // 2], gbs=2 * M * 64 * 4 / (sorted(ts)[len(ts) // 2] * 1e-6) / 1e9)
print(json.dumps(res))
