"""Forward accuracy of one lower-level layer against fp64, for the three ways this repo can compute it and for plain
fp32 torch (what the reference does): max / mean-signed error of t and y relative to the tensor's scale, and how many
ReLU masks differ from the fp64 ones (a flipped mask is what moves a gradient by more than rounding).
usage (GPU): python tools/fwd_err_diag.py [rows]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bignn_b200 import ops
from tests.test_gpu_fused_stack import random_block_graph, run_layer, DEV

torch.backends.cuda.matmul.allow_tf32 = False
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
torch.manual_seed(1)
csr = random_block_graph(rows, 7)
X = torch.randn(rows, 64, device=DEV)
W1 = torch.randn(64, 64, device=DEV) / 8
W2 = torch.randn(64, 64, device=DEV) / 8
b1, b2 = torch.randn(64, device=DEV) * 0.1, torch.randn(64, device=DEV) * 0.1
crp = torch.as_tensor(np.asarray([0, rows], np.int32)).to(DEV)
z = ops.spmm(csr, X, ops.SPMM_GIN, 1.0)
z64 = z.double()
t64p = z64 @ W1.double().t() + b1.double()
t64 = torch.relu(t64p)


def report(name, T, Y, t_in64=None):
    # y is judged against the fp64 transform of THIS path's own t (the error of one transform, not the chain)
    tin = T.double()
    y64p = tin @ W2.double().t() + b2.double()
    y64 = torch.relu(y64p)
    et, ey = (T.double() - t64), (Y.double() - y64)
    st, sy = float(t64.abs().max()), float(y64.abs().max())
    print('%-34s t: max %.2e mean %+.2e | y: max %.2e mean %+.2e | masks != fp64: t %d  y %d  of %d' % (
        name, float(et.abs().max()) / st, float(et.mean()) / st, float(ey.abs().max()) / sy, float(ey.mean()) / sy,
        int(((T > 0) != (t64p > 0)).sum()), int(((Y > 0) != (y64p > 0)).sum()), T.numel()))


Yf, Zf, Tf, _ = run_layer(csr, X, 64, W1, b1, W2, b2, 1, 1, crp, stats=False)
assert torch.equal(Zf, z)
report('fused layer kernel', Tf, Yf)
Tt = ops.gemm_tc(z, W1, True, b1, 1)
report('k_gemm_tc (3xTF32, rotating acc)', Tt, ops.gemm_tc(Tt, W2, True, b2, 1))
Ts = ops.gemm(z, W1, False, True, b1, 1)
report('k_gemm_f32 (fp32 FMA)', Ts, ops.gemm(Ts, W2, False, True, b2, 1))
Tp = torch.relu(z @ W1.t() + b1)
report('torch fp32 (cuBLAS, no TF32)', Tp, torch.relu(Tp @ W2.t() + b2))
Tc = torch.relu(z.cpu() @ W1.cpu().t() + b1.cpu())
report('torch fp32 on the CPU', Tc.to(DEV), torch.relu(Tc @ W2.cpu().t() + b2.cpu()).to(DEV))
