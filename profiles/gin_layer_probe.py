"""Times bignn_gin_layer_fwd on a molecule-like merged graph (default 6 M rows x 64, > L2) against the separate
kernels it replaces.  BIGNN_GL_DEBUG=1 (no gathers) / 2 (no stores) attribute the time to the kernel's roles."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bignn_b200 as B
from bignn_b200 import ops, _lib, synthetic as S
B._lib.load()
rows_target = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
G = rows_target // 30
atom_ptr, nbr_ptr, nbr_idx, _ = S.molecule_graphs(G, 30.0, seed=3)
A = int(atom_ptr[-1])
col = nbr_idx.astype(np.int64) + np.repeat(atom_ptr[:-1].astype(np.int64), np.diff(atom_ptr))[np.repeat(np.arange(A), np.diff(nbr_ptr))]
dev = 'cuda:0'
csr = ops.CSR(torch.as_tensor(nbr_ptr).to(dev), torch.as_tensor(col.astype(np.int32)).to(dev), A)
chunk = atom_ptr[::128].astype(np.int64); chunk = np.append(chunk[chunk < A], A) if chunk[-1] != A else chunk
crp = torch.as_tensor(chunk.astype(np.int32)).to(dev); Sn = crp.numel() - 1
n_tiles = (A + 127) // 128
tile0 = torch.as_tensor(np.clip(np.searchsorted(chunk, np.arange(n_tiles) * 128, side='right') - 1, 0, Sn - 1).astype(np.int32)).to(dev)
pos = torch.as_tensor(np.minimum(np.arange(n_tiles + 1, dtype=np.int64) * 128, A)).to(dev)
tile_edge = csr.row_ptr.index_select(0, pos).contiguous()
X = torch.randn(A, 64, device=dev)
W1 = torch.randn(64, 64, device=dev) / 8; W2 = torch.randn(64, 64, device=dev) / 8
b1 = torch.randn(64, device=dev) * .1; b2 = torch.randn(64, device=dev) * .1
fm = torch.randn(Sn, 64, device=dev) * .1; fa = torch.rand(Sn, 64, device=dev) + .5; fb = torch.randn(64, device=dev) * .1
Y = torch.empty(A, 64, device=dev); Z = torch.empty(A, 64, device=dev); T = torch.empty(A, 64, device=dev)
parts = torch.empty(_lib.call('bignn_gin_layer_stat_records', A, Sn), 2, 64, dtype=torch.float64, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)

def timeit(fn, reps=8):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(0.)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))

def fused(keep, fold, stats):
    _lib.call('bignn_gin_layer_fwd', A, 64, 64, csr.row_ptr, csr.col_idx, csr.nnz, tile_edge, X, X.stride(0),
              fm if fold else None, fa if fold else None, fb if fold else None, crp, Sn, tile0, 1.0, W1, b1, W2, b2, 1, 1,
              Z if keep else None, 64, T if keep else None, 64, Y, 64, parts if stats else None)

def unfused():
    z = ops.spmm(csr, X, ops.SPMM_GIN, 1.0, out=Z)
    t = ops.gemm_tc(z, W1, True, b1, 1, out=T)
    ops.gemm_tc(t, W2, True, b2, 1, out=Y)

res = {'rows': A, 'nnz': csr.nnz, 'chunks': Sn, 'dbg': os.environ.get('BIGNN_GL_DEBUG', '0')}
base = 4.0 * 64 * A
idx = 4.0 * csr.nnz + 4.0 * (A + 1)
for name, keep, fold, stats in (('fused y only', False, False, False), ('fused y + stats + fold', False, True, True),
                                ('fused keep z,t + stats + fold', True, True, True)):
    ms = timeit(lambda: fused(keep, fold, stats))
    nbytes = base * (2 + 2 * keep) + idx
    res[name] = dict(ms=round(ms, 4), algorithmic_GBps=round(nbytes / ms / 1e6, 1))
ms = timeit(unfused)
res['unfused spmm + 2 gemm_tc (no BatchNorm passes)'] = dict(ms=round(ms, 4))
print(json.dumps(res))
