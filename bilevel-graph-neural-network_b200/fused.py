"""The lower-level GIN stack + readout as fused launches (SURVEY 8b `gin_layer_fwd`).

model/layers.py:42-57 per layer is GINConv (aggregate, Linear, act, Linear) -> act -> BatchNorm1d, and
model/layers_aggregation.py:27-42 pools every layer's BatchNorm output.  Here one layer is ONE kernel
(`bignn_gin_layer_fwd`: aggregation + both transforms on the tensor cores + the BatchNorm statistics from the
epilogue) plus a [chunks, 64] finalize; the BatchNorm *apply* is never a pass over the atoms: its affine is
folded into the next layer's aggregation and into the readout (`bignn_readout_fold_fwd`).  The backward is
written out by hand over the path's kernels (segmented BatchNorm backward with the activation derivative
folded in, tensor-core weight gradients, masked backward-input transform, the symmetric SpMM).

Used by BiGNNEngine when every lower layer is a GIN with BatchNorm, a fusable activation and 64 output
channels; everything else goes through the layer-by-layer path (layers.py), which stays the parity reference
for this one (tests/test_gpu_fused_stack.py).
"""
import torch

from . import _lib, ops


class StackSpec(object):
    """What the fused stack needs to know about the model and the merged graph (built once per engine)."""

    def __init__(self, layers, agg, merged, dst_row=None, out_rows=None, bn_stats_sink=None):
        self.layers = list(layers)
        self.merged = merged
        self.style = ops.READOUT_CODES[agg.style]
        self.multi = bool(agg.concat_multi_scale)
        self.dst_row, self.out_rows = dst_row, out_rows
        self.sink = bn_stats_sink
        self.keep_acts = False          # tests: keep the BatchNorm outputs of every layer (materialised on demand)

    def params(self):
        out = []
        for l in self.layers:
            lin1, lin2 = l.conv.nn[0], l.conv.nn[2]
            out += [lin1.weight, lin1.bias, lin2.weight, lin2.bias, l.bn.weight, l.bn.bias]
        return out


def stack_supported(layers, agg, num_node_feat):
    """True when the fused kernels cover this lower level exactly."""
    if agg.style not in ops.READOUT_CODES:
        return False
    din = num_node_feat
    for l in layers:
        if getattr(l, 'type', None) != 'gin' or not l.bn or l.normalize or l.act.code is None:
            return False
        lin1, lin2 = l.conv.nn[0], l.conv.nn[2]
        if lin1.bias is None or lin2.bias is None:
            return False
        if lin1.in_features != din or lin1.out_features != 64 or lin2.out_features != 64:
            return False
        if not _lib.call('bignn_gin_layer_supported', int(din), 64):
            return False
        din = 64
    return len(layers) > 0


class _GinStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, training, x0, *params):
        m = spec.merged
        plan = m.fused_plan()
        A, S, dev = m.A, m.S, x0.device
        _lib.require_device(x0)
        grad_on = training and any(ctx.needs_input_grad)      # (autograd is off inside forward: ask the ctx)
        X, fm, fa, fbeta = x0, None, None, None
        saved, levels = [], []
        L = len(spec.layers)
        for li, layer in enumerate(spec.layers):
            W1, b1, W2, b2, gamma, beta = params[6 * li:6 * li + 6]
            din = W1.shape[1]
            din_pad = (din + 3) // 4 * 4
            a_out = layer.act.code
            a_in = a_out                                   # GINConv's inner act is the same module (layers.py:26-31)
            Y = torch.empty((A, 64), dtype=torch.float32, device=dev)
            Z = torch.empty((A, din_pad), dtype=torch.float32, device=dev) if grad_on else None
            T = torch.empty((A, 64), dtype=torch.float32, device=dev) if grad_on else None
            parts = torch.empty((plan['records'], 2, 64), dtype=torch.float64, device=dev) if training else None
            bn = layer.bn
            _lib.call('bignn_gin_layer_fwd', A, int(din), 64, m.row_ptr, m.col_idx, m.E, plan['tile_edge_ptr'], X,
                      X.stride(0), fm, fa, fbeta,
                      m.chunk_row_ptr, S, plan['tile_chunk0'], 1.0 + layer._eps_value(), W1.contiguous(), b1, W2.contiguous(),
                      b2, int(a_in), int(a_out), Z, Z.stride(0) if Z is not None else 0, T,
                      T.stride(0) if T is not None else 0, Y, Y.stride(0), parts)
            if training:
                mean = torch.empty((S, 64), dtype=torch.float32, device=dev)
                rstd = torch.empty((S, 64), dtype=torch.float32, device=dev)
                stats = torch.empty((2, S, 64), dtype=torch.float64, device=dev)
                fa = torch.empty((S, 64), dtype=torch.float32, device=dev)
                _lib.call('bignn_gin_bn_finalize', parts, m.chunk_row_ptr, S, 64, float(bn.eps), gamma, mean, rstd,
                          stats, fa)
                fm = mean
                if spec.sink is not None:
                    spec.sink.append((bn, stats))          # sharded chunks: the engine replays the running updates
                else:
                    ops.bn_running_update(stats, m.chunk_row_ptr, S, bn.running_mean, bn.running_var,
                                          bn.num_batches_tracked, bn.momentum)
            else:
                # eval: running statistics, the same affine for every chunk
                fa = (gamma * torch.rsqrt(bn.running_var + bn.eps)).unsqueeze(0).expand(S, 64).contiguous()
                fm = bn.running_mean.unsqueeze(0).expand(S, 64).contiguous()
                mean = rstd = None
            fbeta = beta
            saved.append((Y, Z, T, mean, rstd, int(a_in), int(a_out), int(din)))
            if spec.multi or li == L - 1:
                levels.append((li, Y, fm, fa, fbeta))
            X = Y
        nl = len(levels)
        out_rows = spec.out_rows if spec.out_rows is not None else m.G
        alloc = torch.zeros if spec.dst_row is not None else torch.empty
        out = alloc((out_rows, nl * 64), dtype=torch.float32, device=dev)
        for k, (li, Y, mu, a, b) in enumerate(levels):
            _lib.call('bignn_readout_fold_fwd', Y, Y.stride(0), m.seg_ptr, m.G, 64, spec.style, spec.dst_row, mu, a, b,
                      plan['graph_chunk'], out, out.stride(0), k * 64)
        if spec.keep_acts:
            spec.last_levels = levels
        ctx.spec, ctx.saved, ctx.params = spec, saved, params
        ctx.level_of = {lv[0]: k for k, lv in enumerate(levels)}
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, saved, params = ctx.spec, ctx.saved, ctx.params
        m = spec.merged
        A, S, dev = m.A, m.S, dout.device
        dout = ops._f32c(dout)
        grads = [None] * len(params)
        g_next = None                    # gradient w.r.t. this layer's BatchNorm output coming from the layer above
        parts = ops.bn_parts(S, A)
        wsb = _lib.call('bignn_bn_workspace_bytes', S, 64, parts)
        for li in range(len(spec.layers) - 1, -1, -1):
            Y, Z, T, mean, rstd, a_in, a_out, din = saved[li]
            W1, b1, W2, b2, gamma, beta = params[6 * li:6 * li + 6]
            k = ctx.level_of.get(li)
            if g_next is None:
                g = torch.empty((A, 64), dtype=torch.float32, device=dev)
                _lib.call('bignn_readout_bwd', dout, dout.stride(0), k * 64, spec.dst_row, m.seg_ptr, m.G, 64,
                          spec.style, g, g.stride(0), 0)
            else:
                g = g_next
                if k is not None:                          # readout gradient added in place (no separate add pass)
                    _lib.call('bignn_readout_bwd', dout, dout.stride(0), k * 64, spec.dst_row, m.seg_ptr, m.G, 64,
                              spec.style, g, g.stride(0), 1)
            # BatchNorm backward with the outer activation's derivative folded in: dY = d loss / d (Lin2 output)
            dY = torch.empty_like(Y)
            dgamma = torch.empty(64, dtype=torch.float32, device=dev)
            dbeta = torch.empty(64, dtype=torch.float32, device=dev)
            ws = ops._ws(wsb, dev)
            _lib.call('bignn_bn_seg_bwd', Y, Y.stride(0), g, g.stride(0), dY, dY.stride(0), m.chunk_row_ptr, S, 64,
                      parts, gamma, mean, rstd, dgamma, dbeta, a_out, ws, int(wsb))
            del g
            dW2, db2 = ops.linear_bwd_params(dY, T, 'oi', True)
            dT = ops.linear_bwd_input(dY, W2, 'oi', T, a_in)          # (dY W2) * act'(t)
            del dY
            dW1, db1 = ops.linear_bwd_params(dT, Z, 'oi', True)
            if dW1.shape[1] != din:
                dW1 = dW1[:, :din].contiguous()
            grads[6 * li:6 * li + 6] = [dW1, db1, dW2, db2, dgamma, dbeta]
            if li > 0:
                dZ = ops.linear_bwd_input(dT, W1, 'oi', None, 0)
                del dT
                g_next = ops.spmm(m.csr, dZ, ops.SPMM_GIN, 1.0 + spec.layers[li]._eps_value())
                del dZ
        ctx.saved = None
        return (None, None, None) + tuple(grads)


def gin_stack(spec, x0, training):
    """pooled [out_rows, levels*64] of the whole lower level over `spec.merged` (x0 = its feature rows)."""
    return _GinStack.apply(spec, bool(training), x0, *spec.params())


def level_activations(spec):
    """BatchNorm outputs of the pooled levels of the last forward (tests only: one elementwise pass each)."""
    m = spec.merged
    outs = []
    for li, Y, mu, a, b in spec.last_levels:
        seg = torch.repeat_interleave(torch.arange(m.S, device=Y.device),
                                      (m.chunk_row_ptr[1:] - m.chunk_row_ptr[:-1]).long())
        outs.append((Y - mu[seg]) * a[seg] + b)
    return outs
