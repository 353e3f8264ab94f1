"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

Pure-PyTorch restatement of the torch-scatter==1.1.2 entry points the reference
calls (reference Dockerfile:33; call sites model/layers_aggregation.py:3,17,19,54,94,
src/merged_graph.py:2,84, model/layers_meta.py:3,94,144).  Semantics follow the
published 1.1.2 API: the output is zero-filled (fill_value=0), its size along
`dim` is `dim_size` or index.max()+1, and scatter_mean divides by the per-slot
count clamped to >= 1.  The third-party wheel is absent from /root/reference, so
this is "parity unpinned" for the third-party arithmetic.
"""
import torch


def _gen(src, index, dim, out, dim_size, fill_value):
    dim = dim if dim >= 0 else src.dim() + dim
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = src.size(dim)
        index = index.view(shape).expand_as(src)
    if out is None:
        if dim_size is None:
            dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
        size = list(src.size())
        size[dim] = dim_size
        out = src.new_full(size, fill_value)
    return src, out, index, dim


def scatter_add(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    src, out, index, dim = _gen(src, index, dim, out, dim_size, fill_value)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    out = scatter_add(src, index, dim, out, dim_size, fill_value)
    count = scatter_add(torch.ones_like(src), index, dim, None, out.size(dim))
    return out / count.clamp(min=1)


def scatter_max(src, index, dim=-1, out=None, dim_size=None, fill_value=None):
    """Returns (max, argmax).  Empty slots hold `fill_value` (1.1.2 default: the
    dtype's lowest value is used internally and replaced by 0 afterwards)."""
    dim_ = dim if dim >= 0 else src.dim() + dim
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim_] = src.size(dim_)
        index_e = index.view(shape).expand_as(src)
    else:
        index_e = index
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim_] = dim_size
    low = torch.finfo(src.dtype).min if src.is_floating_point() else torch.iinfo(src.dtype).min
    res = src.new_full(size, low)
    res = res.scatter_reduce(dim_, index_e, src, reduce='amax', include_self=True)
    # argmax: first position attaining the max
    hit = (src == res.gather(dim_, index_e))
    pos = torch.arange(src.size(dim_), device=src.device)
    shape = [1] * src.dim()
    shape[dim_] = src.size(dim_)
    pos = pos.view(shape).expand_as(src)
    big = src.size(dim_)
    cand = torch.where(hit, pos, torch.full_like(pos, big))
    arg = torch.full(size, big, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim_, index_e, cand, reduce='amin', include_self=True)
    empty = res == low
    res = res.masked_fill(empty, 0 if fill_value is None else fill_value)
    arg = arg.masked_fill(arg == big, -1)
    return res, arg
