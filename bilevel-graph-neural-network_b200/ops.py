"""Operators of the Bi-GNN path: thin torch-tensor wrappers over the C-ABI plus the
`torch.autograd.Function`s that give them a backward.  torch is plumbing here
(device memory, streams, autograd tape); every arithmetic step is a kernel of
`csrc/` reached through `_lib.call`.
"""
import torch

from . import _lib

ACT_CODES = {'identity': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3}
SPMM_SUM, SPMM_GIN, SPMM_GCN = 0, 1, 2
READOUT_CODES = {'sum': 0, 'avg_pool': 1}


def act_code(name):
    if name not in ACT_CODES:
        raise ValueError('Unknown activation function {}'.format(name))
    return ACT_CODES[name]


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError('bignn_b200 computes in fp32, got {}'.format(t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


LONG_ROW_SEG = int(__import__('os').environ.get('BIGNN_SEG', 32))      # neighbours per work item: a sub-warp walks its item 4 neighbours at a time, so the
                       # item length bounds the dependent-latency chain (8 steps); hub rows become many items


BIG_ITEMS = 64         # = BIGNN_SPMM_BIG_ITEMS (include/bignn_b200.h)


class RowPlan(object):
    """Work-item split of a skewed CSR (host-built once for a static graph)."""

    def __init__(self, row_ptr_host, device, seg=LONG_ROW_SEG):
        import numpy as np
        deg = np.diff(np.asarray(row_ptr_host, np.int64))
        items = np.maximum(1, -(-deg // seg))
        ptr = np.concatenate([[0], np.cumsum(items)])
        self.seg = int(seg)
        self.n_items = int(ptr[-1])
        multi = np.nonzero(items > 1)[0]
        # hub rows (more than BIG_ITEMS items) go last: bignn_spmm_planned_rows_f32 sums them with a whole CTA
        big = items[multi] > BIG_ITEMS
        multi = np.concatenate([multi[~big], multi[big]])
        self.n_big = int(big.sum())
        self.n_multi = int(multi.shape[0])
        as_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(device)
        self.item_ptr = as_dev(ptr)
        self.item_row = as_dev(np.repeat(np.arange(deg.shape[0]), items))
        self.multi_rows = as_dev(multi) if self.n_multi else None


class CSR(object):
    """int32 CSR of a symmetric graph on the device (+ cached GCN deg^-1/2, + an optional
    long-row work-item plan when the host knows the degree distribution is skewed)."""

    def __init__(self, row_ptr, col_idx, n_rows, row_ptr_host=None):
        self.row_ptr, self.col_idx, self.n_rows = row_ptr, col_idx, int(n_rows)
        self._dinv = None
        self.plan = None
        self.row_offset = None      # set for one rank's rows of a row-partitioned graph (graph.PartitionedInteractionGraph)
        self._row_ptr_host = row_ptr_host
        if row_ptr_host is not None and len(row_ptr_host) > 1:
            import numpy as np
            if int(np.diff(np.asarray(row_ptr_host, np.int64)).max()) > LONG_ROW_SEG:
                self.plan = RowPlan(row_ptr_host, row_ptr.device)

    @property
    def nnz(self):
        return int(self.col_idx.numel())

    def ensure_plan(self):
        """a work-item plan for kernels that always run over items (GAT): one item per row when the
        graph has no long rows."""
        if self.plan is None:
            rp = self.row_ptr.cpu().numpy() if self._row_ptr_host is None else self._row_ptr_host
            self.plan = RowPlan(rp, self.row_ptr.device)
        return self.plan

    def dinv(self):
        if self._dinv is None:
            d = torch.empty(self.n_rows, dtype=torch.float32, device=self.row_ptr.device)
            _lib.call('bignn_gcn_dinv', self.row_ptr, self.col_idx, self.n_rows, d)
            self._dinv = d
        return self._dinv


# ----------------------------------------------------------------------------- raw ops
def spmm(csr, x, mode, self_coef=0.0, dinv=None, bias=None, act=0, out=None):
    """y = A x over the rows of `csr`.  For a row-partitioned graph (`csr.row_offset` set) x covers ALL
    nodes in the gathered position space and y only this rank's rows."""
    x = _f32c(x)
    _lib.require_device(x, csr.row_ptr)
    n, d = csr.n_rows, x.shape[1]
    y = out if out is not None else torch.empty((n, d), dtype=torch.float32, device=x.device)
    pl = csr.plan
    roff = csr.row_offset
    if pl is not None and d % 4 == 0 and d <= 512:
        wsb = _lib.call('bignn_spmm_planned_workspace_bytes', pl.n_items, d) if pl.n_multi else 0
        ws = _ws(wsb, x.device) if wsb else None
        _lib.call('bignn_spmm_planned_rows_f32', csr.row_ptr, csr.col_idx, pl.item_ptr, pl.item_row, pl.n_items,
                  pl.seg, pl.multi_rows, pl.n_multi, pl.n_big, x, x.stride(0), y, y.stride(0), n,
                  int(roff) if roff is not None else 0, d, int(mode), float(self_coef), dinv, bias, int(act), ws,
                  int(wsb))
        return y
    if roff is None:
        _lib.call('bignn_spmm_f32', csr.row_ptr, csr.col_idx, x, x.stride(0), y, y.stride(0), n, d,
                  int(mode), float(self_coef), dinv, bias, int(act))
    else:
        _lib.call('bignn_spmm_rows_f32', csr.row_ptr, csr.col_idx, x, x.stride(0), y, y.stride(0), n, int(roff), d,
                  int(mode), float(self_coef), dinv, bias, int(act))
    return y


def gemm(a, b, ta=False, tb=False, bias=None, act=0, out=None):
    """C = act(op(a) op(b) + bias);  op(a) is [M,K], op(b) is [K,N]."""
    a, b = _f32c(a), _f32c(b)
    _lib.require_device(a, b)
    M, K = (a.shape[1], a.shape[0]) if ta else (a.shape[0], a.shape[1])
    K2, N = (b.shape[1], b.shape[0]) if tb else (b.shape[0], b.shape[1])
    if K != K2:
        raise ValueError('gemm: inner dimensions differ ({} vs {})'.format(K, K2))
    c = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=a.device)
    wsb = _lib.call('bignn_gemm_workspace_bytes', M, N, K, int(ta))
    ws = _ws(wsb, a.device) if wsb > 0 else None
    _lib.call('bignn_gemm_f32', int(ta), int(tb), M, N, K, a, a.stride(0), b, b.stride(0),
              c, c.stride(0), bias, int(act), ws, int(wsb))
    return c


import os as _os

TC_MIN_ROWS = int(_os.environ.get('BIGNN_TC_MIN_ROWS', 8192))     # below this the dense transforms run as fp32 FMA (SIMT): a 128-row tensor-core tile grid cannot fill the SMs, and the
                       # small upper-level graphs of the reference datasets (1 309 / 3 242 drugs) then keep exact fp32 products (forward gates 1e-5)
_NO_TC = bool(_os.environ.get('BIGNN_NO_TC'))      # debugging switch: route the transforms to the SIMT kernels


def gemm_tc(a, b, b_is_nk, bias=None, act=0, out=None, mask_y=None, mask_act=0):
    """C = act(a @ op(b) + bias) [* mask_act'(mask_y)] on the tensor cores (tcgen05, 3xTF32)."""
    a, b = _f32c(a), _f32c(b)
    _lib.require_device(a, b)
    M, K = a.shape
    N = b.shape[0] if b_is_nk else b.shape[1]
    if (b.shape[1] if b_is_nk else b.shape[0]) != K:
        raise ValueError('gemm_tc: inner dimensions differ')
    c = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=a.device)
    if mask_y is not None and mask_act:
        mask_y = _f32c(mask_y)
        if tuple(mask_y.shape) != (M, N):
            raise ValueError('gemm_tc: mask shape {} != output shape {}'.format(tuple(mask_y.shape), (M, N)))
        _lib.call('bignn_gemm_tc_masked_f32', M, N, K, a, a.stride(0), b, b.stride(0), int(bool(b_is_nk)), c,
                  c.stride(0), bias, int(act), mask_y, mask_y.stride(0), int(mask_act))
        return c
    _lib.call('bignn_gemm_tc_f32', M, N, K, a, a.stride(0), b, b.stride(0), int(bool(b_is_nk)), c, c.stride(0), bias,
              int(act))
    return c


def use_tc(M, N, K):
    return (not _NO_TC) and M >= TC_MIN_ROWS and N <= 128 and K % 4 == 0 and (K <= 64 or (N <= 64 and K <= 96))


def dw_tc(p, q, colsum_of=-1):
    """(p^T q, column sums of p or q) on the tensor cores; p [M,Np], q [M,Nq]."""
    p, q = _f32c(p), _f32c(q)
    _lib.require_device(p, q)
    M, Np, Nq = p.shape[0], p.shape[1], q.shape[1]
    d = torch.empty((Np, Nq), dtype=torch.float32, device=p.device)
    cs = torch.empty(Nq if colsum_of == 1 else Np, dtype=torch.float32, device=p.device) if colsum_of >= 0 else None
    wsb = _lib.call('bignn_dw_tc_workspace_bytes', M, Np, Nq)
    ws = _ws(wsb, p.device)
    _lib.call('bignn_dw_tc_f32', M, Np, Nq, p, p.stride(0), q, q.stride(0), d, int(colsum_of), cs, ws, int(wsb))
    return d, cs


def use_dw_tc(M, Np, Nq):
    return (not _NO_TC) and M >= TC_MIN_ROWS and Np <= 64 and Nq <= 64 and Np % 4 == 0 and Nq % 4 == 0


def colsum(x):
    x = _f32c(x)
    _lib.require_device(x)
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    wsb = _lib.call('bignn_colsum_workspace_bytes', rows, cols)
    ws = _ws(wsb, x.device)
    _lib.call('bignn_colsum_f32', x, x.stride(0), rows, cols, out, ws, int(wsb))
    return out


def act_bwd(y, dy, act):
    if act == 0:
        return dy
    y, dy = _f32c(y), _f32c(dy)
    dx = torch.empty_like(dy)
    _lib.call('bignn_act_bwd_f32', y, dy, dx, dy.numel(), int(act))
    return dx


def bn_parts(S, rows):
    """Row parts per segment so that S*parts CTAs cover the SMs a few times."""
    if S <= 0:
        return 1
    # streaming-sized inputs (> 64 MB at 64 channels): fill all 8 resident CTAs per SM; small inputs keep fewer,
    # larger parts (their finalize kernels walk the parts and are latency-bound)
    per_sm = 8 if rows >= 262144 else 4
    want = max(1, (148 * per_sm) // S)
    by_rows = max(1, (rows // max(S, 1)) // 64)
    return int(max(1, min(want, by_rows, 256)))


# ----------------------------------------------------------------------------- autograd
class _GinAggregate(torch.autograd.Function):
    """z_i = (1+eps) x_i + sum_{j in N(i)} x_j  (PyG GINConv aggregation, App. A.1).
    The graph is symmetric, so the backward is the same SpMM on the gradient."""

    @staticmethod
    def forward(ctx, x, csr, eps):
        ctx.csr, ctx.eps = csr, float(eps)
        return spmm(csr, x, SPMM_GIN, 1.0 + float(eps))

    @staticmethod
    def backward(ctx, dz):
        return spmm(ctx.csr, dz, SPMM_GIN, 1.0 + ctx.eps), None, None


class _GcnPropagate(torch.autograd.Function):
    """u = act(D^-1/2 (A+I) D^-1/2 h + bias)  (PyG GCNConv propagate + model/layers.py:55)."""

    @staticmethod
    def forward(ctx, h, bias, csr, act, act_bwd_by_consumer=False):
        u = spmm(csr, h, SPMM_GCN, 0.0, csr.dinv(), bias, act)
        ctx.csr, ctx.act, ctx.has_bias = csr, (0 if act_bwd_by_consumer else act), bias is not None
        ctx.save_for_backward(u)
        return u

    @staticmethod
    def backward(ctx, du):
        (u,) = ctx.saved_tensors
        g = act_bwd(u, du, ctx.act)
        dbias = colsum(g) if ctx.has_bias and ctx.needs_input_grad[1] else None
        dh = spmm(ctx.csr, g, SPMM_GCN, 0.0, ctx.csr.dinv(), None, 0) if ctx.needs_input_grad[0] else None
        return dh, dbias, None, None, None


class _SumRows(torch.autograd.Function):
    """out_i = sum_{e in rows(i)} x_e over an arbitrary (non-symmetric) CSR; used to add the
    per-entry row gradients of the pair decoder per drug, deterministically."""

    @staticmethod
    def forward(ctx, x, csr):
        return spmm(csr, x, SPMM_SUM)

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError


class _LinearAct(torch.autograd.Function):
    """y = act(x W^T + b) (layout 'oi', nn.Linear) or act(x W + b) (layout 'io', PyG)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, layout, act_bwd_by_consumer=False, input_act=0):
        N = weight.shape[0] if layout == 'oi' else weight.shape[1]
        if use_tc(x.shape[0], N, x.shape[1]):
            y = gemm_tc(x, weight, layout == 'oi', bias, act)
        else:
            y = gemm(x, weight, False, layout == 'oi', bias, act)
        # act_bwd_by_consumer: the BatchNorm that consumes y folds act'(y) into its own backward
        # (bignn_bn_seg_bwd input_act), so the incoming gradient is already w.r.t. the pre-activation
        # input_act: x is the output of that activation and ITS producer was told act_bwd_by_consumer: dX is
        # returned w.r.t. the activation's input (mask fused into the backward-input GEMM's epilogue)
        ctx.act, ctx.layout, ctx.input_act = (0 if act_bwd_by_consumer else act), layout, int(input_act)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        g = act_bwd(y, dy, ctx.act)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_input(g, weight, ctx.layout, x if ctx.input_act else None, ctx.input_act)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = linear_bwd_params(g, x, ctx.layout, ctx.needs_input_grad[2], ctx.needs_input_grad[1])
        return dx, dw, db, None, None, None, None


def linear_bwd_input(g, weight, layout, x_mask=None, input_act=0):
    """dX = g W ('oi') / g W^T ('io'), times act'(x_mask) when the producer of x left its activation's derivative to
    this consumer (mask fused into the tensor-core kernel's epilogue)."""
    K = weight.shape[1] if layout == 'oi' else weight.shape[0]
    if use_tc(g.shape[0], K, g.shape[1]):
        # oi: W is [N,K] = op(B)^T stored [K',N'] -> b_is_nk False; io: W is [K,N] -> b_is_nk True
        return gemm_tc(g, weight, layout == 'io', mask_y=x_mask if input_act else None, mask_act=input_act)
    dx = gemm(g, weight, False, layout == 'io')
    if input_act:
        dx = act_bwd(x_mask, dx, input_act)
    return dx


def linear_bwd_params(g, x, layout, want_b=True, want_w=True):
    """(dW, db) of y = x W^T + b ('oi': dW = g^T x) / y = x W + b ('io': dW = x^T g)."""
    dw = db = None
    if want_w and use_dw_tc(g.shape[0], g.shape[1], x.shape[1]):
        if layout == 'oi':
            dw, db = dw_tc(g, x, 0 if want_b else -1)
        else:
            dw, db = dw_tc(x, g, 1 if want_b else -1)
    else:
        if want_w:
            dw = gemm(g, x, True, False) if layout == 'oi' else gemm(x, g, True, False)
        if want_b:
            db = colsum(g)
    return dw, db


class _SegBatchNorm(torch.autograd.Function):
    """Train-mode BatchNorm1d with independent statistics per row segment."""

    @staticmethod
    def forward(ctx, x, gamma, beta, seg_row_ptr, S, running_mean, running_var, nbt, eps, momentum, stats_out,
                input_act=0):
        x = _f32c(x)
        _lib.require_device(x)
        rows, C = x.shape
        parts = bn_parts(S, rows)
        y = torch.empty_like(x)
        mean = torch.empty((S, C), dtype=torch.float32, device=x.device)
        rstd = torch.empty((S, C), dtype=torch.float32, device=x.device)
        wsb = _lib.call('bignn_bn_workspace_bytes', S, C, parts)
        ws = _ws(wsb, x.device)
        _lib.call('bignn_bn_seg_fwd', x, x.stride(0), y, y.stride(0), seg_row_ptr, S, C, parts, gamma, beta,
                  float(eps), float(momentum), running_mean, running_var, nbt, mean, rstd, stats_out, ws, int(wsb))
        ctx.S, ctx.parts, ctx.seg, ctx.input_act = S, parts, seg_row_ptr, int(input_act)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = _f32c(dy)
        rows, C = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        wsb = _lib.call('bignn_bn_workspace_bytes', ctx.S, C, ctx.parts)
        ws = _ws(wsb, x.device)
        _lib.call('bignn_bn_seg_bwd', x, x.stride(0), dy, dy.stride(0), dx, dx.stride(0), ctx.seg, ctx.S, C,
                  ctx.parts, gamma, mean, rstd, dgamma, dbeta, ctx.input_act, ws, int(wsb))
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None, None


class _Readout(torch.autograd.Function):
    """Multi-scale segment readout: out[dst_row[g], l*D:(l+1)*D] = pool_l(acts_l[seg g])."""

    @staticmethod
    def forward(ctx, seg_ptr, G, style, dst_row, out_rows, *acts):
        acts = [_f32c(a) for a in acts]
        _lib.require_device(*acts)
        D = acts[0].shape[1]
        L = len(acts)
        alloc = torch.zeros if dst_row is not None else torch.empty
        out = alloc((out_rows, L * D), dtype=torch.float32, device=acts[0].device)
        for l, a in enumerate(acts):
            _lib.call('bignn_readout_fwd', a, a.stride(0), seg_ptr, G, D, style, dst_row, out, out.stride(0), l * D)
        ctx.seg, ctx.G, ctx.style, ctx.dst_row, ctx.D = seg_ptr, G, style, dst_row, D
        ctx.rows = [a.shape[0] for a in acts]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _f32c(dout)
        grads = []
        for l, rows in enumerate(ctx.rows):
            if not ctx.needs_input_grad[5 + l]:
                grads.append(None)
                continue
            dx = torch.empty((rows, ctx.D), dtype=torch.float32, device=dout.device)
            _lib.call('bignn_readout_bwd', dout, dout.stride(0), l * ctx.D, ctx.dst_row, ctx.seg, ctx.G, ctx.D,
                      ctx.style, dx, dx.stride(0), 0)
            grads.append(dx)
        return (None, None, None, None, None) + tuple(grads)


class _GatedReadout(torch.autograd.Function):
    """out[g] = sum over the atoms of graph g of sigmoid(gate) * weight (the gated / "attention" readout of
    model/layers_aggregation.py:90-94), product and segment sum in one launch each way."""

    @staticmethod
    def forward(ctx, gate, w, seg_ptr, G):
        gate, w = _f32c(gate), _f32c(w)
        _lib.require_device(gate, w)
        D = w.shape[1]
        out = torch.empty((G, D), dtype=torch.float32, device=w.device)
        _lib.call('bignn_readout_gated_fwd', gate, gate.stride(0), w, w.stride(0), seg_ptr, int(G), D, out, out.stride(0))
        ctx.seg, ctx.G = seg_ptr, int(G)
        ctx.save_for_backward(gate, w)
        return out

    @staticmethod
    def backward(ctx, dout):
        gate, w = ctx.saved_tensors
        dout = _f32c(dout)
        dg, dw = torch.empty_like(gate), torch.empty_like(w)
        _lib.call('bignn_readout_gated_bwd', gate, gate.stride(0), w, w.stride(0), dout, dout.stride(0), ctx.seg, ctx.G,
                  w.shape[1], dg, dg.stride(0), dw, dw.stride(0))
        return dg, dw, None, None


def gated_readout(gate, w, seg_ptr, G):
    if w.shape[1] % 4 != 0 or not w.is_cuda:
        return readout([gate_mul(gate, w)], seg_ptr, G, 'sum')
    return _GatedReadout.apply(gate, w, seg_ptr, G)


class _PairGatherNorm(torch.autograd.Function):
    """z[p] = [normalize(h[ids[p,0]]) || normalize(h[ids[p,1]])]; the backward adds the
    per-entry gradients per drug through the entry CSR (deterministic, no atomics)."""

    @staticmethod
    def forward(ctx, h, ids, entry_csr):
        h = _f32c(h)
        _lib.require_device(h, ids)
        P, D = ids.shape[0], h.shape[1]
        z = torch.empty((P, 2 * D), dtype=torch.float32, device=h.device)
        nrm = torch.empty((P, 2), dtype=torch.float32, device=h.device)
        _lib.call('bignn_pair_gather_norm_fwd', h, h.stride(0), ids, P, D, z, z.stride(0), nrm)
        ctx.entry_csr = entry_csr
        ctx.save_for_backward(h, ids, nrm)
        return z

    @staticmethod
    def backward(ctx, dz):
        h, ids, nrm = ctx.saved_tensors
        dz = _f32c(dz)
        P, D = ids.shape[0], h.shape[1]
        drows = torch.empty((2 * P, D), dtype=torch.float32, device=h.device)
        _lib.call('bignn_pair_gather_norm_bwd', h, h.stride(0), ids, P, D, dz, dz.stride(0), nrm, drows,
                  drows.stride(0))
        dh = spmm(ctx.entry_csr, drows, SPMM_SUM)
        return dh, None, None


_PD_WS = {}          # device -> persistent zero-initialised workspace of the fused pair decoder (holds its counter)


def _pair_decoder_ws(nbytes, device):
    ws = _PD_WS.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _PD_WS[device] = ws
    return ws


def pair_decoder_supported(D, widths):
    """widths = output widths of the scorer's Linear layers (2 or 3 of them)."""
    if len(widths) not in (2, 3):
        return False
    w = list(widths) + [0] * (3 - len(widths))
    return bool(_lib.call('bignn_pair_decoder_supported', int(D), len(widths), int(w[0]), int(w[1]), int(w[2])))


class _PairDecoder(torch.autograd.Function):
    """(scores, loss) = fused gather + normalise + concat + MLP + head + loss; see bignn_pair_decoder_fwd.
    args: h, ids, entry_csr, head, y (float) or labels (int32) or None, then W0, b0, W1, b1[, W2, b2]."""

    @staticmethod
    def forward(ctx, h, ids, entry_csr, head, target, *params):
        h = _f32c(h)
        _lib.require_device(h, ids)
        nl = len(params) // 2
        Ws = [p.contiguous() for p in params[0::2]]
        bs = list(params[1::2])
        P, D = ids.shape[0], h.shape[1]
        n = [w.shape[0] for w in Ws] + [0] * (3 - nl)
        n_out = Ws[-1].shape[0]
        dev = h.device
        scores = torch.empty((P, n_out), dtype=torch.float32, device=dev)
        nrm = torch.empty((P, 2), dtype=torch.float32, device=dev)
        h1 = torch.empty((P, n[0]), dtype=torch.float32, device=dev)
        h2 = torch.empty((P, n[1]), dtype=torch.float32, device=dev) if nl == 3 else None
        loss = torch.empty((), dtype=torch.float32, device=dev) if target is not None else None
        wsb = _lib.call('bignn_pair_decoder_workspace_bytes', P, D, nl, n[0], n[1], n[2])
        ws = _pair_decoder_ws(wsb, dev)
        y = target if (target is not None and head != 2) else None
        labels = target if (target is not None and head == 2) else None
        _lib.call('bignn_pair_decoder_fwd', h, h.stride(0), ids, P, D, nl, Ws[0], bs[0], n[0], Ws[1], bs[1], n[1],
                  Ws[2] if nl == 3 else None, bs[2] if nl == 3 else None, n[2], int(head), y, labels, scores,
                  scores.stride(0), nrm, h1, h2, loss, ws, int(ws.numel()))
        ctx.entry_csr, ctx.head, ctx.nl, ctx.n, ctx.has_b = entry_csr, int(head), nl, n, [b is not None for b in bs]
        ctx.save_for_backward(h, ids, target, scores, nrm, h1, h2, *Ws)
        ctx.mark_non_differentiable(scores)
        if loss is None:
            return scores, scores.new_zeros(())
        return scores, loss

    @staticmethod
    def backward(ctx, _dscores, dloss):
        h, ids, target, scores, nrm, h1, h2, *Ws = ctx.saved_tensors
        nl, n, head = ctx.nl, ctx.n, ctx.head
        P, D = ids.shape[0], h.shape[1]
        dev = h.device
        dloss = _f32c(dloss)
        drows = torch.empty((2 * P, D), dtype=torch.float32, device=dev)
        dWs = [torch.empty_like(w) for w in Ws]
        dbs = [torch.empty(w.shape[0], dtype=torch.float32, device=dev) for w in Ws]
        wsb = _lib.call('bignn_pair_decoder_workspace_bytes', P, D, nl, n[0], n[1], n[2])
        ws = _pair_decoder_ws(wsb, dev)
        y = target if head != 2 else None
        labels = target if head == 2 else None
        _lib.call('bignn_pair_decoder_bwd', h, h.stride(0), ids, P, D, nl, Ws[0], n[0], Ws[1], n[1],
                  Ws[2] if nl == 3 else None, n[2], head, y, labels, scores, scores.stride(0), nrm, h1, h2, dloss, drows,
                  drows.stride(0), dWs[0], dbs[0], dWs[1], dbs[1], dWs[2] if nl == 3 else None,
                  dbs[2] if nl == 3 else None, ws, int(ws.numel()))
        dh = spmm(ctx.entry_csr, drows, SPMM_SUM)
        grads = []
        for l in range(nl):
            grads += [dWs[l], dbs[l] if ctx.has_b[l] else None]
        return (dh, None, None, None, None) + tuple(grads)


def pair_decoder(h, ids, entry_csr, head, target, params):
    """Returns (scores [P, n_out], loss scalar).  head: 0 sigmoid+BCE, 1 logits+BCEWithLogits, 2 logits+CE."""
    return _PairDecoder.apply(h, ids, entry_csr, int(head), target, *params)


class _BCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, y, logits):
        pred = _f32c(pred).view(-1)
        y = _f32c(y).view(-1)
        _lib.require_device(pred, y)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        _lib.call('bignn_bce_logits_fwd' if logits else 'bignn_bce_fwd', pred, y, pred.numel(), loss)
        ctx.logits = logits
        ctx.save_for_backward(pred, y)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        pred, y = ctx.saved_tensors
        dloss = _f32c(dloss)
        dp = torch.empty_like(pred)
        _lib.call('bignn_bce_logits_bwd' if ctx.logits else 'bignn_bce_bwd', pred, y, pred.numel(), dloss, dp)
        return dp, None, None


class _GatConv(torch.autograd.Function):
    """PyG 1.1.2 GATConv propagate, 1 head: out = edge_softmax-weighted sum of h + bias."""

    @staticmethod
    def forward(ctx, h, att, bias, csr, slope, group_target, n_block=None):
        h = _f32c(h)
        _lib.require_device(h, att)
        n, D = h.shape
        n_block = n if n_block is None else int(n_block)
        att = att.contiguous()
        bias = bias.contiguous() if bias is not None else None
        out = torch.empty_like(h)
        scratch = torch.empty(4 * max(n, 1), dtype=torch.float32, device=h.device)
        a = att.reshape(-1)
        pl = csr.ensure_plan()
        wsb = _lib.call('bignn_gat_fwd_workspace_bytes', pl.n_items, D)
        ws = _ws(wsb, h.device)
        _lib.call('bignn_gat_fwd', csr.row_ptr, csr.col_idx, pl.item_ptr, pl.item_row, pl.n_items, pl.seg,
                  pl.multi_rows, pl.n_multi, n, n_block, D, h, h.stride(0), a, bias, float(slope),
                  int(group_target), out, out.stride(0), scratch, ws, int(wsb))
        ctx.csr, ctx.slope, ctx.group, ctx.n_block = csr, float(slope), int(group_target), n_block
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.att_shape = att.shape
        ctx.save_for_backward(h, a, bias, out, scratch)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, a, bias, out, scratch = ctx.saved_tensors
        dout = _f32c(dout)
        n, D = h.shape
        dh = torch.empty_like(h)
        dpq = torch.empty(2 * max(n, 1), dtype=torch.float32, device=h.device)
        pl = ctx.csr.ensure_plan()
        wsb = _lib.call('bignn_gat_bwd_workspace_bytes', n, D, pl.n_items)
        ws = _ws(wsb, h.device)
        _lib.call('bignn_gat_bwd', ctx.csr.row_ptr, ctx.csr.col_idx, pl.item_ptr, pl.item_row, pl.n_items, pl.seg,
                  pl.multi_rows, pl.n_multi, n, ctx.n_block, D, h, h.stride(0), a, bias, ctx.slope,
                  ctx.group, out, out.stride(0), dout, dout.stride(0), scratch, dh, dh.stride(0), dpq, ws, int(wsb))
        nb, T = ctx.n_block, n // ctx.n_block
        datt = dbias = None
        if ctx.needs_input_grad[1]:
            v = dpq[:2 * n].view(2, n)
            parts = [gemm(v[:, t * nb:(t + 1) * nb], h[t * nb:(t + 1) * nb]).reshape(1, 2 * D) for t in range(T)]
            datt = (parts[0] if T == 1 else torch.cat(parts, 0)).reshape(ctx.att_shape)
        if bias is not None and ctx.needs_input_grad[2]:
            parts = [colsum(dout[t * nb:(t + 1) * nb]) for t in range(T)]
            dbias = (parts[0] if T == 1 else torch.stack(parts, 0)).reshape(ctx.bias_shape)
        return dh, datt, dbias, None, None, None, None


class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        x = _f32c(x)
        _lib.require_device(x)
        y = torch.empty_like(x)
        _lib.call('bignn_act_fwd_f32', x, y, x.numel(), int(act))
        ctx.act = act
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return act_bwd(y, dy, ctx.act), None


class _PReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        x = _f32c(x)
        _lib.require_device(x, weight)
        y = torch.empty_like(x)
        _lib.call('bignn_prelu_fwd_f32', x, y, x.shape[0], x.shape[1], weight, weight.numel())
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        t = torch.empty_like(x)
        _lib.call('bignn_prelu_bwd_f32', x, dy, dx, t, x.shape[0], x.shape[1], weight, weight.numel())
        dw = colsum(t) if weight.numel() > 1 else colsum(t.view(-1, 1))
        return dx, dw


class _RowNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        _lib.require_device(x)
        y = torch.empty_like(x)
        nrm = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        _lib.call('bignn_rownorm_fwd_f32', x, x.stride(0), y, y.stride(0), x.shape[0], x.shape[1], nrm)
        ctx.save_for_backward(y, nrm)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(y)
        _lib.call('bignn_rownorm_bwd_f32', y, y.stride(0), dy, dy.stride(0), nrm, dx, dx.stride(0), y.shape[0],
                  y.shape[1])
        return dx


class _GateMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, w):
        gate, w = _f32c(gate), _f32c(w)
        _lib.require_device(gate, w)
        o = torch.empty_like(w)
        _lib.call('bignn_gate_mul_fwd_f32', gate, w, o, w.numel())
        ctx.save_for_backward(gate, w)
        return o

    @staticmethod
    def backward(ctx, do):
        gate, w = ctx.saved_tensors
        do = _f32c(do)
        dg, dw = torch.empty_like(gate), torch.empty_like(w)
        _lib.call('bignn_gate_mul_bwd_f32', gate, w, do, dg, dw, w.numel())
        return dg, dw


class _PairDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, act):
        z = _f32c(z)
        _lib.require_device(z)
        P, D = z.shape[0], z.shape[1] // 2
        out = torch.empty(P, dtype=torch.float32, device=z.device)
        _lib.call('bignn_pair_dot_fwd_f32', z, z.stride(0), P, D, out, int(act))
        ctx.act = act
        ctx.save_for_backward(z, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, out = ctx.saved_tensors
        dout = _f32c(dout)
        dz = torch.empty_like(z)
        _lib.call('bignn_pair_dot_bwd_f32', z, z.stride(0), z.shape[0], z.shape[1] // 2, out, dout, int(ctx.act), dz,
                  dz.stride(0))
        return dz, None


class _CE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        logits = _f32c(logits)
        _lib.require_device(logits, labels)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        _lib.call('bignn_ce_fwd', logits, logits.stride(0), labels, logits.shape[0], logits.shape[1], loss)
        ctx.save_for_backward(logits, labels)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, labels = ctx.saved_tensors
        d = torch.empty_like(logits)
        _lib.call('bignn_ce_bwd', logits, logits.stride(0), labels, logits.shape[0], logits.shape[1], _f32c(dloss), d,
                  d.stride(0))
        return d, None


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32c(a), _f32c(b)
        _lib.require_device(a, b)
        o = torch.empty_like(a)
        _lib.call('bignn_add_f32', a, b, o, a.numel())
        return o

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return _Add.apply(a, b)


def gat_conv(h, att, bias, csr, negative_slope=0.2, group='source', n_block=None):
    """n_block: several edge types batched as a block-diagonal graph (h [T*n_block, D], att [T, 2D],
    bias [T, D]); None = one graph."""
    if group not in ('source', 'target'):
        raise ValueError('GAT softmax group must be source or target')
    return _GatConv.apply(h, att, bias, csr, negative_slope, 1 if group == 'target' else 0, n_block)


class _StackedLinear(torch.autograd.Function):
    """H[t*N:(t+1)*N] = x @ W_t for every edge type t (PyG weight layout [in, out]): the node
    transforms of NodeModelAggrByEdge written into one stacked buffer for the batched GAT."""

    @staticmethod
    def forward(ctx, x, *weights):
        x = _f32c(x)
        N, T, D = x.shape[0], len(weights), weights[0].shape[1]
        h = torch.empty((T * N, D), dtype=torch.float32, device=x.device)
        for t, w in enumerate(weights):
            dst = h[t * N:(t + 1) * N]
            if use_tc(N, D, x.shape[1]):
                gemm_tc(x, w, False, out=dst)
            else:
                gemm(x, w, False, False, out=dst)
        ctx.save_for_backward(x, *weights)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, *weights = ctx.saved_tensors
        dh = _f32c(dh)
        N = x.shape[0]
        dx = None
        dws = []
        for t, w in enumerate(weights):
            g = dh[t * N:(t + 1) * N]
            if ctx.needs_input_grad[0]:
                part = gemm_tc(g, w, True) if use_tc(N, x.shape[1], g.shape[1]) else gemm(g, w, False, True)
                dx = part if dx is None else add(dx, part)
            if ctx.needs_input_grad[1 + t]:
                dws.append(dw_tc(x, g, -1)[0] if use_dw_tc(N, x.shape[1], g.shape[1]) else gemm(x, g, True, False))
            else:
                dws.append(None)
        return (dx,) + tuple(dws)


def stacked_linear(x, weights):
    return _StackedLinear.apply(x, *weights)


def activation(x, act):
    return x if act == 0 else _Act.apply(x, act)


def prelu(x, weight):
    return _PReLU.apply(x, weight)


def row_normalize(x):
    return _RowNorm.apply(x)


def gate_mul(gate, w):
    return _GateMul.apply(gate, w)


def pair_dot(z, act=0):
    return _PairDot.apply(z, act)


def cross_entropy(logits, labels_i32):
    return _CE.apply(logits, labels_i32)


def gin_aggregate(x, csr, eps=0.0):
    return _GinAggregate.apply(x, csr, eps)


def gcn_propagate(h, bias, csr, act=0, act_bwd_by_consumer=False):
    return _GcnPropagate.apply(h, bias, csr, act, act_bwd_by_consumer)


def linear_act(x, weight, bias=None, act=0, layout='oi', act_bwd_by_consumer=False, input_act=0):
    return _LinearAct.apply(x, weight, bias, act, layout, act_bwd_by_consumer, input_act)


def seg_batch_norm(x, gamma, beta, seg_row_ptr, S, running_mean, running_var, nbt, eps=1e-5, momentum=0.1,
                   stats_out=None, input_act=0):
    """stats_out: optional fp64 [2,S,C] buffer receiving the per-segment mean / unbiased variance
    (multi-GPU: running buffers are then advanced by `bn_running_update` after an exchange).
    input_act: x is the output of this activation and its producer was told `act_bwd_by_consumer`: the
    backward returns the gradient w.r.t. the activation's input (one pass less over [rows, C])."""
    return _SegBatchNorm.apply(x, gamma, beta, seg_row_ptr, S, running_mean, running_var, nbt, eps, momentum,
                               stats_out, input_act)


def bn_running_update(seg_stats, seg_row_ptr, S, running_mean, running_var, nbt, momentum=0.1):
    C = running_mean.numel()
    _lib.call('bignn_bn_running_update', seg_stats, seg_row_ptr, int(S), int(C), float(momentum), running_mean,
              running_var, nbt)


def bn_eval(x, gamma, beta, running_mean, running_var, eps=1e-5):
    x = _f32c(x)
    _lib.require_device(x)
    y = torch.empty_like(x)
    _lib.call('bignn_bn_eval_fwd', x, x.stride(0), y, y.stride(0), x.shape[0], x.shape[1], gamma, beta,
              float(eps), running_mean, running_var)
    return y


def readout(acts, seg_ptr, G, style='avg_pool', dst_row=None, out_rows=None):
    if style not in READOUT_CODES:
        raise NotImplementedError('{} is not implemented'.format(style))
    return _Readout.apply(seg_ptr, int(G), READOUT_CODES[style], dst_row,
                          int(out_rows if out_rows is not None else G), *acts)


def pair_gather_norm(h, ids, entry_csr):
    return _PairGatherNorm.apply(h, ids, entry_csr)


def bce(pred, y, logits=False):
    return _BCE.apply(pred, y, logits)
