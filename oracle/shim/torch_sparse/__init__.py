"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Restatement of torch-sparse==0.2.4 `coalesce` (reference Dockerfile:34; call site
model/layers_util.py:6,161): sort the COO entries by the key row*n+col, drop
duplicate keys, and reduce duplicate values with scatter_<op>.
"""
import torch
import torch_scatter


def coalesce(index, value, m, n, op='add', fill_value=0):
    row, col = index
    if row.numel() == 0:
        return index, value
    unique, inv = torch.unique(row * n + col, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype, device=inv.device)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    index = torch.stack([row[perm], col[perm]], dim=0)
    if value is not None:
        fn = getattr(torch_scatter, 'scatter_{}'.format(op))
        value = fn(value, inv, 0, None, perm.size(0), fill_value) if op != 'max' \
            else fn(value, inv, 0, None, perm.size(0))
        if isinstance(value, tuple):
            value = value[0]
    return index, value
