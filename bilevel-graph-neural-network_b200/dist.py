"""Multi-GPU plumbing of the path (one process per GPU, torch.distributed; NCCL over NVLink on
the B200 box, gloo in the CPU tests).

Partitioning: the all-drug lower pass is sharded BY DRUG at chunk granularity --
whole 128-graph chunks of src/train.py:62-71 stay on one rank, so lower-level BatchNorm needs no
communication and keeps the reference's per-chunk statistics.  The exchange step is the pooled
drug embeddings: every rank writes its rows of init_x[N, L*D] and the ranks sum the disjoint
row sets (an all-gather expressed as an all-reduce of zero-padded buffers, which NVSwitch reduces
in-switch); its backward is the identity because the upper level is evaluated on the full
init_x by every rank.  Lower-level weight gradients are partial sums over a rank's chunks and are
all-reduced once per step in one flat buffer; BatchNorm running buffers replay the reference's
sequential per-chunk momentum updates from all-gathered per-chunk statistics.
"""
import numpy as np
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_chunks(chunk_weights, world):
    """Contiguous blocks of chunks per rank, balanced by weight (atoms).  Returns [(lo, hi)]."""
    w = np.asarray(chunk_weights, np.float64)
    n = len(w)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target))
        # closest prefix boundary, monotone, and leaving at least nothing negative
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, n)] - target):
            i -= 1
        i = max(i, bounds[-1])
        i = min(i, n)
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class _SumDisjointRows(torch.autograd.Function):
    """init_x = sum over ranks of zero-padded row blocks (each drug row is written by exactly one
    rank).  Backward: identity -- every rank evaluates the upper level on the full matrix, so
    d loss / d (own rows) is simply the matching rows of its own d init_x."""

    @staticmethod
    def forward(ctx, x, group):
        x = x.contiguous()
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


def sum_disjoint_rows(x, group=None):
    if not is_dist():
        return x
    return _SumDisjointRows.apply(x, group)


def all_reduce_grads(params, group=None):
    """One flat SUM all-reduce over the gradients of `params` (lower-level weights: each rank
    holds the partial sum over its own chunks)."""
    if not is_dist():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def gather_chunk_stats(stats_local, s_max, group=None):
    """all-gather of per-chunk BatchNorm statistics [2, S_local, C] (padded to s_max chunks)
    -> [world, 2, s_max, C]."""
    world = dist.get_world_size(group)
    two, s_loc, C = stats_local.shape
    pad = torch.zeros((2, s_max, C), dtype=stats_local.dtype, device=stats_local.device)
    pad[:, :s_loc] = stats_local
    out = torch.empty((world, 2, s_max, C), dtype=stats_local.dtype, device=stats_local.device)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return out
