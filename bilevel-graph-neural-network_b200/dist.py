"""Multi-GPU plumbing of the path (one process per GPU, torch.distributed; NCCL over NVLink on
the B200 box, gloo in the CPU tests).

Partitioning: the all-drug lower pass is sharded BY DRUG at chunk granularity --
whole 128-graph chunks of src/train.py:62-71 stay on one rank, so lower-level BatchNorm needs no
communication and keeps the reference's per-chunk statistics.  The exchange step is an ALL-GATHER of
the pooled drug embeddings (every rank contributes the rows of its shard, every rank receives
init_x[N, L*D]); its backward takes the rank's own rows (replicated upper level) or REDUCE-SCATTERs
the partial gradients (row-partitioned upper level).  Lower-level weight gradients are partial sums over a rank's chunks and are
all-reduced once per step in one flat buffer; BatchNorm running buffers replay the reference's
sequential per-chunk momentum updates from all-gathered per-chunk statistics.
"""
import numpy as np
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


# ---- optional per-collective device timing (bench.py --dist-profile; eager steps only, never under capture)
TIMING = None          # None = off; else {tag: [(start event, end event), ...]}


class _timed(object):
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if TIMING is not None and torch.cuda.is_available():
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if TIMING is not None and torch.cuda.is_available():
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMING.setdefault(self.tag, []).append((self.e0, e1))


def timing_summary(steps=1):
    """{tag: (ms per step, calls per step)} of the collectives recorded since TIMING was set (includes the
    wait for the slowest rank)."""
    torch.cuda.synchronize()
    return {k: (sum(a.elapsed_time(b) for a, b in v) / steps, len(v) // steps) for k, v in (TIMING or {}).items()}


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


def all_reduce_grads(params, group=None, extra=None):
    """One flat SUM all-reduce over the gradients of `params` (lower-level weights: each rank
    holds the partial sum over its own chunks).  `extra`: a small fp32 vector that rides along (the pair-batch
    checksum); its rank sum is returned."""
    if not is_dist():
        return None
    grads = [p.grad for p in params if p.grad is not None]
    if not grads and extra is None:
        return None
    flat = torch.cat([g.reshape(-1) for g in grads] + ([extra.reshape(-1)] if extra is not None else []))
    with _timed('grad_all_reduce'):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    if grads:
        torch._foreach_copy_(grads, views)           # one multi-tensor launch instead of one copy per gradient
    return flat[off:] if extra is not None else None


def gather_chunk_stats(stats_local, s_max, group=None):
    """all-gather of per-chunk BatchNorm statistics [K, S_local, C] (K = 2 per layer; padded to s_max chunks)
    -> [world, K, s_max, C]."""
    world = dist.get_world_size(group)
    K, s_loc, C = stats_local.shape
    pad = torch.zeros((K, s_max, C), dtype=stats_local.dtype, device=stats_local.device)
    pad[:, :s_loc] = stats_local
    out = torch.empty((world, K, s_max, C), dtype=stats_local.dtype, device=stats_local.device)
    with _timed('bn_chunk_stats_all_gather'):
        dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return out


# ------------------------------------------------------------------------------------------------
# Row-partitioned upper level: interaction-graph rows (= edges by source drug) are split
# over the ranks (graph.RowPartition); per layer: local GEMM on own rows -> ALL-GATHER of the
# transformed rows -> local SpMM on own rows -> BatchNorm over all N rows through an all-reduce of
# (sum x, sum x^2).  The graph is symmetric, so each backward is the same exchange on the gradient.
def _group_size(group):
    return dist.get_world_size(group) if is_dist() else 1


def gather_rows(x_loc, pg):
    """[n_loc, D] rows of this rank -> [world*n_max, D] rows of all ranks in the position space of
    graph.RowPartition (equal-count all-gather, in place in the output buffer: NCCL's in-place form,
    sendbuff = recvbuff + rank*count)."""
    D = x_loc.shape[1]
    out = torch.empty((pg.n_pad, D), dtype=x_loc.dtype, device=x_loc.device)
    mine = out[pg.row_offset:pg.row_offset + pg.n_max]
    mine[:pg.n_loc].copy_(x_loc)
    if pg.n_loc < pg.n_max:
        mine[pg.n_loc:].zero_()
    if pg.world > 1:
        src = mine if x_loc.is_cuda else mine.clone()          # gloo (CPU tests): no aliasing of in/out
        with _timed('upper_rows_all_gather'):
            dist.all_gather_into_tensor(out.view(-1), src.reshape(-1), group=pg.group)
    return out


class PooledLayout(object):
    """Where the pooled drug rows of every rank live in the all-gathered [world, n_max, L*D] buffer: rank r's block
    holds the drugs of its lower-level shard (ascending drug row), `perm[g]` = position of drug row g."""

    def __init__(self, rank, world, n_max, n_loc, perm, group=None):
        self.rank, self.world, self.n_max, self.n_loc, self.perm, self.group = rank, world, n_max, n_loc, perm, group


class _GatherPooledRows(torch.autograd.Function):
    """The exchange step between the levels: ALL-GATHER of the pooled drug embeddings -- every rank
    contributes the rows of its lower-level shard, every rank receives init_x [N, L*D].
    Backward: with a replicated upper level every rank holds the whole d init_x and simply takes the rows of its own
    block; with a row-partitioned upper level a rank holds d init_x only for its own interaction-graph rows, and the
    blocks are summed over the ranks with a REDUCE-SCATTER (half the bytes of the all-reduce this replaces)."""

    @staticmethod
    def forward(ctx, blk, lay, partial_grad):
        n_max, world = lay.n_max, lay.world
        D = blk.shape[1]
        buf = torch.empty((world * n_max, D), dtype=blk.dtype, device=blk.device)
        mine = buf[lay.rank * n_max:(lay.rank + 1) * n_max]
        mine.copy_(blk[:n_max])
        src = mine if blk.is_cuda else mine.clone()            # gloo (CPU tests): no aliasing of in/out
        with _timed('pooled_rows_all_gather'):
            dist.all_gather_into_tensor(buf.view(-1), src.reshape(-1), group=lay.group)
        ctx.lay, ctx.partial, ctx.rows = lay, bool(partial_grad), blk.shape[0]
        return buf.index_select(0, lay.perm)

    @staticmethod
    def backward(ctx, g):
        lay = ctx.lay
        n_max, world = lay.n_max, lay.world
        D = g.shape[1]
        scat = torch.zeros((world * n_max, D), dtype=g.dtype, device=g.device)
        scat.index_copy_(0, lay.perm, g.contiguous())
        if ctx.partial:
            out = torch.empty((n_max, D), dtype=g.dtype, device=g.device)
            with _timed('pooled_rows_reduce_scatter'):
                if g.is_cuda:
                    dist.reduce_scatter_tensor(out.view(-1), scat.view(-1), op=dist.ReduceOp.SUM, group=lay.group)
                else:                                          # gloo has no reduce-scatter
                    dist.all_reduce(scat, op=dist.ReduceOp.SUM, group=lay.group)
                    out.copy_(scat[lay.rank * n_max:(lay.rank + 1) * n_max])
        else:
            out = scat[lay.rank * n_max:(lay.rank + 1) * n_max]
        gb = torch.zeros((ctx.rows, D), dtype=g.dtype, device=g.device)
        gb[:n_max] = out
        return gb, None, None


def gather_pooled_rows(blk, lay, partial_grad):
    return _GatherPooledRows.apply(blk, lay, partial_grad)


class _GcnPropagateRows(torch.autograd.Function):
    """u_loc = act(rows_loc(D^-1/2 (A+I) D^-1/2) all_gather(h_loc) + bias); backward: the same
    all-gather + local SpMM on act'(u) * du (A symmetric); dbias is this rank's partial column sum."""

    @staticmethod
    def forward(ctx, h_loc, bias, pg, act, act_bwd_by_consumer=False):
        from . import ops
        H = gather_rows(h_loc, pg)
        u = ops.spmm(pg.csr, H, ops.SPMM_GCN, 0.0, pg.csr.dinv(), bias, act)
        ctx.pg, ctx.act, ctx.has_bias = pg, (0 if act_bwd_by_consumer else act), bias is not None
        ctx.save_for_backward(u)
        return u

    @staticmethod
    def backward(ctx, du):
        from . import ops
        (u,) = ctx.saved_tensors
        pg = ctx.pg
        g = ops.act_bwd(u, du, ctx.act)
        dbias = ops.colsum(g) if ctx.has_bias and ctx.needs_input_grad[1] else None
        dh = None
        if ctx.needs_input_grad[0]:
            G = gather_rows(g, pg)
            dh = ops.spmm(pg.csr, G, ops.SPMM_GCN, 0.0, pg.csr.dinv(), None, 0)
        return dh, dbias, None, None, None


def gcn_propagate_rows(h_loc, bias, pg, act=0, act_bwd_by_consumer=False):
    return _GcnPropagateRows.apply(h_loc, bias, pg, act, act_bwd_by_consumer)


class _RowsBatchNorm(torch.autograd.Function):
    """Train-mode BatchNorm1d whose batch is the rows of all ranks: local fp64 sums, one [2, C] fp64
    all-reduce, local apply (bignn_bn_rows_*).  dgamma / dbeta come out rank-summed (global)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, pg, running_mean, running_var, nbt, eps, momentum, input_act=0):
        from . import ops, _lib
        x = ops._f32c(x)
        _lib.require_device(x)
        rows, C = x.shape
        parts = ops.bn_parts(1, max(rows, 1))
        wsb = _lib.call('bignn_bn_rows_workspace_bytes', C, parts)
        ws = ops._ws(wsb, x.device)
        sums = torch.empty((2, C), dtype=torch.float64, device=x.device)
        _lib.call('bignn_bn_rows_sums', x, x.stride(0), None, 0, rows, C, parts, None, None, sums, ws, int(wsb))
        if pg.world > 1:
            with _timed('upper_bn_stats_all_reduce'):
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=pg.group)
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        rstd = torch.empty(C, dtype=torch.float32, device=x.device)
        _lib.call('bignn_bn_rows_fwd_apply', x, x.stride(0), y, y.stride(0), rows, C, parts, sums, int(pg.n),
                  gamma, beta, float(eps), float(momentum), running_mean, running_var, nbt, mean, rstd)
        ctx.pg, ctx.parts, ctx.input_act = pg, parts, int(input_act)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops, _lib
        x, gamma, mean, rstd = ctx.saved_tensors
        pg, parts = ctx.pg, ctx.parts
        dy = ops._f32c(dy)
        rows, C = x.shape
        wsb = _lib.call('bignn_bn_rows_workspace_bytes', C, parts)
        ws = ops._ws(wsb, x.device)
        sums = torch.empty((2, C), dtype=torch.float64, device=x.device)
        _lib.call('bignn_bn_rows_sums', x, x.stride(0), dy, dy.stride(0), rows, C, parts, mean, rstd, sums, ws,
                  int(wsb))
        if pg.world > 1:
            with _timed('upper_bn_stats_all_reduce'):
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=pg.group)
        dx = torch.empty_like(x)
        _lib.call('bignn_bn_rows_bwd_apply', x, x.stride(0), dy, dy.stride(0), dx, dx.stride(0), rows, C, parts,
                  gamma, mean, rstd, sums, int(pg.n), ctx.input_act)
        dbeta, dgamma = sums[0].to(torch.float32), sums[1].to(torch.float32)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None


def rows_batch_norm(x, gamma, beta, pg, running_mean, running_var, nbt, eps=1e-5, momentum=0.1, input_act=0):
    return _RowsBatchNorm.apply(x, gamma, beta, pg, running_mean, running_var, nbt, eps, momentum, input_act)


class _GatherRowsReplicatedConsumer(torch.autograd.Function):
    """All-gather of the final upper-level embeddings for a consumer that every rank evaluates
    identically (the pair scorer over the whole pair batch): every rank then holds the same gradient of
    the gathered matrix, so the backward of the gather is this rank's own row block of it."""

    @staticmethod
    def forward(ctx, x_loc, pg):
        ctx.pg = pg
        return gather_rows(x_loc, pg)

    @staticmethod
    def backward(ctx, g):
        pg = ctx.pg
        return g[pg.row_offset:pg.row_offset + pg.n_loc].contiguous(), None


def gather_rows_for_replicated_consumer(x_loc, pg):
    return _GatherRowsReplicatedConsumer.apply(x_loc, pg)
