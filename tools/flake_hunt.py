#!/usr/bin/env python
"""Hunts the one-off deviation of tests/test_gpu_kernels.py::test_spmm_gcn_vs_oracle[49-tanh] (5e-5 instead of 2e-6,
seen twice in ~20 fresh-box runs in round 1).  For every case of that test: the kernel is run R
times on NaN-POISONED output / deg^-1/2 buffers (a read of anything the kernel should have written itself shows up as a
NaN or a changed bit), with a freshly built CSR object every few iterations, and each result is compared bit for bit
with the first one and against the CPU oracle.  Run it in several fresh processes:

    for i in 1 2 3 4 5; do python tools/flake_hunt.py 400; done
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bignn_b200 as B
from bignn_b200 import ops, _lib
from oracle import bignn_oracle as O

R = int(sys.argv[1]) if len(sys.argv) > 1 else 200
DEV = 'cuda:0'
B._lib.load()
gold = os.path.join(ROOT, 'tests', 'golden')
ds = O.PackedDataset.load(os.path.join(gold, 'drugbank_packed.npz'))
n = ds.N
ei = torch.from_numpy(np.stack([ds.ddi_row, ds.ddi_col]))
ptr = np.concatenate([[0], np.cumsum(np.bincount(ds.ddi_row, minlength=n))]).astype(np.int64)
bad = 0
for D, act in [(49, 'tanh'), (64, 'relu'), (64, 'identity'), (320, 'sigmoid')]:
    g = torch.Generator().manual_seed(D)
    h = torch.randn(n, D, generator=g)
    bias = torch.randn(D, generator=g)
    want = O._act(act, O.gcn_conv(h, ei, {'l.conv.weight': torch.eye(D), 'l.conv.bias': bias}, 'l'))
    want2 = O._act(act, O.gcn_conv(h, ei, {'l.conv.weight': torch.eye(D), 'l.conv.bias': bias}, 'l'))
    if not torch.equal(want, want2):
        print('ORACLE not reproducible for', D, act, float((want - want2).abs().max()))
    first = {}                      # per kernel variant (row-sequential / work-item plan): each is deterministic in itself
    worst = 0.0
    for it in range(R):
        if it % 8 == 0:             # a fresh CSR (and a fresh, poisoned deg^-1/2 buffer) every few iterations
            rp = torch.as_tensor(ptr.astype(np.int32)).to(DEV)
            ci = torch.as_tensor(ds.ddi_col.astype(np.int32)).to(DEV)
            planned = it % 16 == 0
            csr = ops.CSR(rp, ci, n, row_ptr_host=ptr if planned else None)
            d = torch.full((n,), float('nan'), device=DEV)
            _lib.call('bignn_gcn_dinv', csr.row_ptr, csr.col_idx, n, d)
            csr._dinv = d
        hd, bd = h.to(DEV), bias.to(DEV)
        out = torch.full((n, D), float('nan'), device=DEV)
        ops.spmm(csr, hd, ops.SPMM_GCN, 0.0, csr.dinv(), bd, ops.act_code(act), out=out)
        got = out.cpu()
        first.setdefault(planned, got)
        err = float((got - want).abs().max() / want.abs().max())
        worst = max(worst, err)
        if not torch.equal(got, first[planned]) or err >= 2e-6 or not bool(torch.isfinite(got).all()):
            bad += 1
            dd = (got - want).abs()
            print('DEVIATION D={} act={} iteration {}: err {:.3g}, bit-equal to first run: {}, finite: {}, worst element {}'
                  .format(D, act, it, err, torch.equal(got, first[planned]), bool(torch.isfinite(got).all()),
                          np.unravel_index(int(dd.argmax()), dd.shape)), flush=True)
    print('D={} act={}: {} runs, worst error vs oracle {:.3g}'.format(D, act, R, worst), flush=True)
print('flake hunt: {} deviating runs'.format(bad))
sys.exit(1 if bad else 0)
