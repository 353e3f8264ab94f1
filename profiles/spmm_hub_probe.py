"""Hub-row combine of the work-item SpMM (bignn_spmm_planned_rows_f32): a degree-skewed interaction graph
(p(v) ~ rank^-0.8, the S-ddi generator) with the hub rows summed by a whole CTA (n_big) vs by one sub-warp
(n_big = 0, the first version).  Prints CUDA-event times and the maximum difference between the two."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
import bignn_b200 as B
from bignn_b200 import ops, synthetic as S
from bignn_b200.graph import InteractionGraph
B._lib.load()
N, M = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
t = time.time()
_, row, col = S.interaction_graph(N, M, 0)
g = InteractionGraph(N, row, col, 'cuda:0')
pl = g.csr.plan
print('graph %.1fs: N %d nnz %d max degree %d items %d multi %d big %d' % (
    time.time() - t, N, g.csr.nnz, int(np.bincount(row).max()), pl.n_items, pl.n_multi, pl.n_big))
x = torch.randn(N, 64, device='cuda:0')


def run():
    return ops.spmm(g.csr, x, ops.SPMM_GCN, 0.0, g.csr.dinv(), None, 0)


def timed(reps=10):
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_new = timed()
y_new = run()
nb = pl.n_big
pl.n_big = 0
t_old = timed()
y_old = run()
pl.n_big = nb
print('ms per SpMM: hub rows by CTA %.4f, by one sub-warp %.4f; gather %.1f GB/s; max |diff| / max |y| %.3g' % (
    t_new, t_old, g.csr.nnz * 256 / t_new / 1e6, float((y_new - y_old).abs().max() / y_old.abs().max())))
