"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the Bi-GNN hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package must never do so.

What it restates (reference file:line, all relative to /root/reference):
  * per-graph edge canonicalisation   model/layers_util.py:148-166 (+ PyG 1.1.2
    to_undirected / torch-sparse 0.2.4 coalesce)
  * block-diagonal merge              src/merged_graph.py:27-96
  * unique-graph first-appearance order, pair labels   src/batch.py:105-144
  * negative sampling (rejection rule, draw order, CPython set order)
                                      src/batch.py:61-103
  * all-drug lower pass chunking      src/train.py:48-72
  * NodeEmbedding gin / gcn / gat     model/layers.py:42-63 (+ PyG 1.1.2 GINConv /
    GCNConv / GATConv)
  * readout avg_pool / sum, multi-scale, init_x row writes
                                      model/layers_aggregation.py:27-42,66-75
  * L2-normalise + gather + MLP + sigmoid pair scorer
                                      model/layers_link_pred.py:34-71, layers_util.py:12-57
  * BCE / BCEWithLogits / CE loss     model/layers.py:79-89
  * one train step                    src/train.py:75-108,177-182

Parity pin: the reference has no tests and its arithmetic lives in un-vendored
third-party wheels (torch-geometric==1.1.2, torch-scatter==1.1.2,
torch-sparse==0.2.4, Dockerfile:32-34).  This port is therefore pinned against
outputs of the reference's OWN Python sources executed in the build container on
top of oracle/shim (oracle/make_golden.py -> tests/golden/*.npz); see
tests/test_oracle_golden.py.  The third-party arithmetic itself stays "parity
unpinned" (restated from the published 1.1.2 API).

Arithmetic is plain torch CPU fp32 (or fp64 when `dtype=torch.float64`), written
as explicit index_add / matmul so that summation order equals the reference's
COO order.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# dataset view
# --------------------------------------------------------------------------
class PackedDataset(object):
    """Flat-array view of what utils/data/dataset.py:47 keeps as networkx objects.

    atom_ptr[N+1], x[sumA,F] (one-hot), nbr_ptr[sumA+1]/nbr_idx[nnz] (per-graph
    local neighbour ids, rows ascending, neighbours ascending), ddi_row/ddi_col
    (directed, sorted train interaction edges over drug rows), pairs dict.
    """

    def __init__(self, d):
        self.gids = np.asarray(d['gids'], np.int64)
        self.atom_ptr = np.asarray(d['atom_ptr'], np.int64)
        self.x = np.asarray(d['x_u8'] if 'x_u8' in d else d['x'], np.float32)
        self.nbr_ptr = np.asarray(d['nbr_ptr'], np.int64)
        self.nbr_idx = np.asarray(d['nbr_idx'], np.int64)
        self.ddi_row = np.asarray(d['ddi_row'], np.int64)
        self.ddi_col = np.asarray(d['ddi_col'], np.int64)
        self.train_pairs = np.asarray(d['train_pairs'], np.int64)
        self.gs_map = {int(g): i for i, g in enumerate(self.gids)}
        self.pairs = {}
        if 'pair_keys' in d:
            for (a, b), l in zip(np.asarray(d['pair_keys']).tolist(),
                                 np.asarray(d['pair_labels']).tolist()):
                self.pairs[(a, b)] = int(l)
        self.num_node_feat = self.x.shape[1]
        self.N = len(self.gids)
        keys = d.files if hasattr(d, 'files') else d.keys()
        self.etypes = {k[len('etype_row/'):]: (np.asarray(d[k], np.int64), np.asarray(d['etype_col/' + k[len('etype_row/'):]], np.int64))
                       for k in sorted(keys) if k.startswith('etype_row/')}

    @classmethod
    def load(cls, path):
        return cls(np.load(path))

    def mol_undirected_edges(self, row):
        """The molecule's undirected bond list [(u,v)], one orientation each --
        what `list(g.edges)` holds before model/layers_util.py:150 sorts it."""
        a0, a1 = self.atom_ptr[row], self.atom_ptr[row + 1]
        n = int(a1 - a0)
        ptr = self.nbr_ptr[a0:a1 + 1]
        cnt = np.diff(ptr)
        src = np.repeat(np.arange(n), cnt)
        dst = self.nbr_idx[ptr[0]:ptr[-1]]
        keep = src < dst
        return np.stack([src[keep], dst[keep]], 1), n

    def look_up_label(self, g1, g2):
        """utils/data/dataset.py:394-403; None when the pair is unknown."""
        l = self.pairs.get((g1, g2))
        if l is None:
            l = self.pairs.get((g2, g1))
        return l


# --------------------------------------------------------------------------
# indexing (bit-exact part)
# --------------------------------------------------------------------------
def coalesce_undirected(edges, n):
    """sorted(list(g.edges)) -> to_undirected -> coalesce  (layers_util.py:148-166).
    Returns int64 [2, 2m'] lexicographically sorted, duplicate-free."""
    if len(edges) == 0:
        return np.zeros((2, 0), np.int64)
    e = np.asarray(sorted(map(tuple, np.asarray(edges).tolist())), np.int64)
    row = np.concatenate([e[:, 0], e[:, 1]])
    col = np.concatenate([e[:, 1], e[:, 0]])
    key = np.unique(row * n + col)
    return np.stack([key // n, key % n]).astype(np.int64)


def unique_graphs_in_order(batch_gids):
    """First-appearance order scanning pairs row-major (src/batch.py:112-113,131-136)."""
    seen = {}
    for g1, g2 in np.asarray(batch_gids).tolist():
        if g1 not in seen:
            seen[g1] = len(seen)
        if g2 not in seen:
            seen[g2] = len(seen)
    return list(seen.keys())


def merge_graphs(ds, gids):
    """src/merged_graph.py:27-96 over the graphs `gids` (already unique, ordered).
    Per-graph conversion is done graph by graph, as the reference does each step."""
    xs, eis, batch = [], [], []
    ind_list, edge_ind_list = [], []
    cn = ce = 0
    for i, gid in enumerate(gids):
        row = ds.gs_map[int(gid)]
        und, n = ds.mol_undirected_edges(row)
        ei = coalesce_undirected(und, n)
        xs.append(ds.x[ds.atom_ptr[row]:ds.atom_ptr[row + 1]])
        eis.append(ei + cn)
        batch.append(np.full((n,), i, np.int64))
        ind_list.append((cn, cn + n))
        edge_ind_list.append((ce, ce + ei.shape[1]))
        cn += n
        ce += ei.shape[1]
    return dict(
        x=np.concatenate(xs, 0), edge_index=np.concatenate(eis, 1),
        batch=np.concatenate(batch), ind_list=np.asarray(ind_list, np.int64),
        edge_ind_list=np.asarray(edge_ind_list, np.int64),
        graph_sizes=np.asarray([b - a for a, b in ind_list], np.int64),
        gids_to_batch_ind={int(g): i for i, g in enumerate(gids)},
        gids=np.asarray(gids, np.int64))


def all_drug_chunks(gids, batch_size):
    """Pair/chunk schedule of the all-drug pass (src/train.py:52-71).
    Returns a list of [P_c,2] gid-pair arrays."""
    gids = list(gids)
    pairs = [(gids[i], gids[i + 1]) for i in range(0, len(gids) - 2, 2)]
    pairs.append((gids[-2], gids[-1]))
    pairs = np.asarray(pairs, np.int64)
    bs = int(len(pairs) / 2) if batch_size * 2 >= len(pairs) else batch_size
    out = []
    i = 0
    for i in range(0, pairs.shape[0] - bs, bs):
        out.append(pairs[i:i + bs])
    out.append(pairs[i + bs:])
    return out


def sample_negative_pairs(ds, positive_gids, sampled_gids, pos_edge_set=None,
                          num_negative_samples=1, rng=np.random):
    """src/batch.py:61-103 with enforce_sampling_amongst_same_graphs=True and
    enforce_negative=True (the shipped defaults, src/config.py).  `rng` must expose
    `choice`; the reference uses the global numpy MT19937 stream.  Result order is
    the iteration order of a CPython set of (gid, gid) tuples, as in the reference."""
    neg = set()
    gid_ind = 0
    n = len(sampled_gids)
    max_pairs = ((n * (n - 1)) / 2) - len(positive_gids)
    gs_map = ds.gs_map
    if pos_edge_set is None:
        pos_edge_set = set(zip(ds.ddi_row.tolist(), ds.ddi_col.tolist()))
    pos_edges = set(pos_edge_set)
    pos_edges.update((gs_map[int(a)], gs_map[int(b)]) for a, b in positive_gids)
    target = min(max_pairs, len(positive_gids) * num_negative_samples)
    while True:
        if len(neg) == target:
            break
        orig = sampled_gids[gid_ind % n]
        cand = rng.choice(sampled_gids, size=1)[0]
        gid_ind += 1
        while ((orig, cand) in neg or (cand, orig) in neg or orig == cand
               or (gs_map[int(orig)], gs_map[int(cand)]) in pos_edges
               or (gs_map[int(cand)], gs_map[int(orig)]) in pos_edges):
            orig = sampled_gids[gid_ind % n]
            gid_ind += 1
            cand = rng.choice(sampled_gids, size=1)[0]
        neg.add((orig, cand))
    order = {k: 0 for k in neg}           # set -> dict -> list, as batch.py:101-103
    return np.asarray(list(order.keys()), np.int64).reshape(-1, 2)


def pair_labels(ds, batch_gids):
    """src/batch.py:117-130: dataset label if the pair is known in either
    orientation (this includes val/test pairs and their pre-made negatives),
    otherwise the sampled negative's 0."""
    out = []
    for g1, g2 in np.asarray(batch_gids).tolist():
        l = ds.look_up_label(g1, g2)
        out.append(0 if l is None else l)
    return np.asarray(out, np.int64)


# --------------------------------------------------------------------------
# layer specs / parameters
# --------------------------------------------------------------------------
def parse_specs(lines):
    """'Name:k=v,k=v' strings (model/layers_factory.py:16-24)."""
    out = []
    for s in lines:
        s = s.strip()
        if not s:
            continue
        sp = s.split(':')
        kv = {}
        if len(sp) > 1:
            for item in sp[1].split(','):
                k = item.split('=')
                kv[k[0]] = '='.join(k[1:])
        out.append((sp[0], kv))
    return out


def _b(s):
    return {'True': True, 'False': False}[s]


def calc_mlp_dims(dim, out_dim=1, division=2):
    """model/layers_link_pred.py:34-41."""
    dims = []
    while dim > out_dim:
        dim = dim // division
        dims.append(dim)
    return dims[:-1]


def init_params(specs, num_node_feat, num_labels=2, interaction_num_node_feat=None,
                seed=0, dtype=torch.float32, num_edge_types=3):
    """Default-initialised parameters/buffers with the reference's state_dict
    names and shapes; init distributions as in App. A."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul(bound).to(dtype)

    def linear(prefix, fin, fout, xavier_gain=None):
        if xavier_gain is None:          # nn.Linear default: kaiming_uniform(a=sqrt 5)
            bound = 1.0 / math.sqrt(fin)
        else:
            bound = xavier_gain * math.sqrt(6.0 / (fin + fout))
        sd[prefix + '.weight'] = uni((fout, fin), bound)
        sd[prefix + '.bias'] = uni((fout,), 1.0 / math.sqrt(fin))

    seen_ne = False
    for i, (name, lf) in enumerate(specs):
        p = 'layers.%d' % i
        if name == 'NodeEmbedding':
            higher = _b(lf['higher_level']) if 'higher_level' in lf else False
            if 'input_dim' in lf:
                fin = int(lf['input_dim'])
            else:
                assert not seen_ne
                fin = interaction_num_node_feat if higher else num_node_feat
            seen_ne = True
            fout = int(lf['output_dim'])
            if lf['type'] == 'gin':
                sd[p + '.conv.eps'] = torch.zeros(1, dtype=dtype)
                linear(p + '.conv.nn.0', fin, fout)
                linear(p + '.conv.nn.2', fout, fout)
            elif lf['type'] == 'gcn':
                sd[p + '.conv.weight'] = uni((fin, fout), math.sqrt(6.0 / (fin + fout)))
                sd[p + '.conv.bias'] = torch.zeros(fout, dtype=dtype)
            elif lf['type'] == 'gat':
                sd[p + '.conv.weight'] = uni((fin, fout), math.sqrt(6.0 / (fin + fout)))
                sd[p + '.conv.att'] = uni((1, 1, 2 * fout), math.sqrt(6.0 / (1 + 2 * fout)))
                sd[p + '.conv.bias'] = torch.zeros(fout, dtype=dtype)
            else:
                raise ValueError(lf['type'])
            if lf['act'] == 'prelu':
                sd[p + '.act.weight'] = torch.full((fout,), 0.25, dtype=dtype)
            if _b(lf['bn']):
                sd[p + '.bn.weight'] = torch.ones(fout, dtype=dtype)
                sd[p + '.bn.bias'] = torch.zeros(fout, dtype=dtype)
                sd[p + '.bn.running_mean'] = torch.zeros(fout, dtype=dtype)
                sd[p + '.bn.running_var'] = torch.ones(fout, dtype=dtype)
                sd[p + '.bn.num_batches_tracked'] = torch.zeros((), dtype=torch.int64)
        elif name == 'MetaLayer':
            # model/layers_meta.py:61-79: one GCNConv / GATConv per edge type except 'none'
            higher = True
            if 'input_dim' in lf:
                fin = int(lf['input_dim'])
            else:
                fin = interaction_num_node_feat
            fout = int(lf['output_dim'])
            kind = lf['node_model'].split('_')[0]
            for e in range(num_edge_types - 1):
                q = p + '.node_model.GNNS.%d' % e
                sd[q + '.weight'] = uni((fin, fout), math.sqrt(6.0 / (fin + fout)))
                if kind == 'gat':
                    sd[q + '.att'] = uni((1, 1, 2 * fout), math.sqrt(6.0 / (1 + 2 * fout)))
                sd[q + '.bias'] = torch.zeros(fout, dtype=dtype)
        elif name == 'LinkPredictor' and lf['type'] == 'mlp_concat':
            d = int(lf['mlp_dim']) * 2
            multi = _b(lf['multi_label_pred']) if lf.get('multi_label_pred') else False
            out = (num_labels + 1) if multi else 1
            hidden = calc_mlp_dims(d, out, 8) if multi else calc_mlp_dims(d, 1, 8)
            ch = [d] + hidden + [out]
            for k in range(len(ch) - 1):
                linear(p + '.mlp_concat.layers.%d' % k, ch[k], ch[k + 1], xavier_gain=math.sqrt(2.0))
    return sd


def state_from_npz(z, prefix, dtype=torch.float32):
    sd = {}
    for k in z.files:
        if k.startswith(prefix):
            v = torch.from_numpy(np.asarray(z[k]))
            sd[k[len(prefix):]] = v.to(dtype) if v.is_floating_point() else v.clone()
    return sd


# --------------------------------------------------------------------------
# arithmetic
# --------------------------------------------------------------------------
def _act(name, x, prelu_w=None):
    if name == 'relu':
        return torch.relu(x)
    if name == 'identity':
        return x
    if name == 'sigmoid':
        return torch.sigmoid(x)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'prelu':
        return F.prelu(x, prelu_w)
    raise ValueError('Unknown activation function {}'.format(name))


def _scatter_rows(src, index, n):
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def gin_conv(x, ei, P, p, act, prelu_w=None):
    """PyG 1.1.2 GINConv (App. A.1): nn((1+eps) x + sum_{j->i} x_j); self loops
    removed; nn = Linear -> act -> Linear sharing the outer act (layers.py:26-31)."""
    row, col = ei
    keep = row != col
    row, col = row[keep], col[keep]
    agg = _scatter_rows(x.index_select(0, row), col, x.shape[0])
    z = (1 + P[p + '.conv.eps']) * x + agg
    h = F.linear(z, P[p + '.conv.nn.0.weight'], P[p + '.conv.nn.0.bias'])
    h = _act(act, h, prelu_w)
    return F.linear(h, P[p + '.conv.nn.2.weight'], P[p + '.conv.nn.2.bias'])


def _with_self_loops(ei, n):
    row, col = ei
    keep = row != col
    loop = torch.arange(n, dtype=torch.long, device=row.device)
    return torch.cat([row[keep], loop]), torch.cat([col[keep], loop])


def gcn_conv(x, ei, P, p):
    """PyG 1.1.2 GCNConv (App. A.2), edge_weight=None."""
    n = x.shape[0]
    h = x @ P[p + '.conv.weight']
    row, col = _with_self_loops(ei, n)
    w = torch.ones(row.shape[0], dtype=h.dtype, device=h.device)
    deg = _scatter_rows(w, row, n)
    dinv = deg.pow(-0.5)
    dinv[dinv == float('inf')] = 0
    norm = dinv[row] * w * dinv[col]
    out = _scatter_rows(norm.view(-1, 1) * h.index_select(0, row), col, n)
    return out + P[p + '.conv.bias']


def gat_conv(x, ei, P, p, softmax_group='source', negative_slope=0.2):
    """PyG 1.1.2 GATConv, 1 head (App. A.3).  softmax_group='source' is the 1.1.x
    grouping (edge_index[0]); 'target' the >=1.2 one.  Graphs here are symmetric,
    the two differ numerically."""
    n = x.shape[0]
    h = x @ P[p + '.conv.weight']
    row, col = _with_self_loops(ei, n)
    att = P[p + '.conv.att'].view(-1)
    d = h.shape[1]
    a = (h.index_select(0, col) * att[:d]).sum(-1) + (h.index_select(0, row) * att[d:]).sum(-1)
    a = F.leaky_relu(a, negative_slope)
    grp = row if softmax_group == 'source' else col
    amax = torch.full((n,), -float('inf'), dtype=a.dtype, device=a.device).scatter_reduce(0, grp, a, 'amax')
    e = (a - amax[grp]).exp()
    s = _scatter_rows(e, grp, n)
    alpha = e / (s[grp] + 1e-16)
    out = _scatter_rows(alpha.view(-1, 1) * h.index_select(0, row), col, n)
    return out + P[p + '.conv.bias']


def batch_norm(x, P, p, training, momentum=0.1, eps=1e-5):
    """torch BatchNorm1d (App. A.8); updates the running buffers in P in place."""
    if not training:
        return F.batch_norm(x, P[p + '.bn.running_mean'], P[p + '.bn.running_var'],
                            P[p + '.bn.weight'], P[p + '.bn.bias'], False, momentum, eps)
    out = F.batch_norm(x, P[p + '.bn.running_mean'], P[p + '.bn.running_var'],
                       P[p + '.bn.weight'], P[p + '.bn.bias'], True, momentum, eps)
    P[p + '.bn.num_batches_tracked'] += 1
    return out


def node_embedding(x, ei, P, p, lf, training, gat_group='source'):
    """model/layers.py:42-63: conv -> act -> bn -> (optional) row L2-normalise.
    Returns (stored_output, returned_output)."""
    act = lf['act']
    pw = P.get(p + '.act.weight')
    t = lf['type']
    if t == 'gin':
        h = gin_conv(x, ei, P, p, act, pw)
    elif t == 'gcn':
        h = gcn_conv(x, ei, P, p)
    elif t == 'gat':
        h = gat_conv(x, ei, P, p, gat_group)
    else:
        raise ValueError('Unknown node embedding layer type {}'.format(t))
    h = _act(act, h, pw)
    if _b(lf['bn']):
        h = batch_norm(h, P, p, training)
    out = F.normalize(h, p=2, dim=1) if _b(lf['normalize']) else h
    return out


def readout(acts, batch, G, style='avg_pool'):
    """torch-scatter 1.1.2 scatter_mean / scatter_add over the batch vector
    (layers_aggregation.py:17-19,34-41, App. A.7)."""
    outs = []
    for a in acts:
        s = _scatter_rows(a, batch, G)
        if style == 'avg_pool':
            c = _scatter_rows(torch.ones_like(a), batch, G)
            s = s / c.clamp(min=1)
        elif style != 'sum':
            raise NotImplementedError('{} is not implemented'.format(style))
        outs.append(s)
    return torch.cat(outs, 1) if len(outs) > 1 else outs[0]


def link_pred(h, ids, P, p, lf, num_labels=2):
    """model/layers_link_pred.py:43-67 (ids = [P,2] rows of `h`)."""
    h = F.normalize(h, p=2, dim=1)
    g1 = h.index_select(0, ids[:, 0])
    g2 = h.index_select(0, ids[:, 1])
    if lf['type'] == 'dot_product':
        return torch.sigmoid((g1 * g2).sum(1))
    multi = _b(lf['multi_label_pred']) if lf.get('multi_label_pred') else False
    z = torch.cat([g1, g2], 1)
    k = 0
    while (p + '.mlp_concat.layers.%d.weight' % (k + 1)) in P:
        z = torch.relu(F.linear(z, P[p + '.mlp_concat.layers.%d.weight' % k],
                                P[p + '.mlp_concat.layers.%d.bias' % k]))
        k += 1
    z = F.linear(z, P[p + '.mlp_concat.layers.%d.weight' % k], P[p + '.mlp_concat.layers.%d.bias' % k])
    return z if multi else torch.sigmoid(z)


def loss_fn(pred, y, kind):
    """model/layers.py:79-89."""
    if kind == 'BCE':
        return F.binary_cross_entropy(pred.view(-1), y.to(pred.dtype))
    if kind == 'BCEWithLogits':
        return F.binary_cross_entropy_with_logits(pred.view(-1), y.to(pred.dtype))
    if kind == 'CE':
        return F.cross_entropy(pred, y.long())
    raise ValueError('Unknown loss layer type {}'.format(kind))


# --------------------------------------------------------------------------
# the path end to end
# --------------------------------------------------------------------------
class OracleModel(object):
    """Sequential layer list driven by the reference's spec strings
    (model/model.py:34-62), over a dict of leaf tensors named like the state_dict."""

    def __init__(self, specs, state, dtype=torch.float32, gat_group='source', device='cpu'):
        self.specs = specs
        self.dtype = dtype
        self.device = torch.device(device)      # 'cuda': the same eager code on the GPU (bench.py gpu_eager_baseline)
        self.gat_group = gat_group
        self.P = {}
        for k, v in state.items():
            v = v.clone().to(self.device)
            if v.is_floating_point():
                v = v.to(dtype)
                if not ('running_' in k or k.endswith('.eps')):
                    v.requires_grad_(True)
            self.P[k] = v
        names = [n for n, _ in specs]
        self.i_agg = names.index('NodeAggregation') if 'NodeAggregation' in names else None
        self.i_load = names.index('LoadInteractionLayer') if 'LoadInteractionLayer' in names else None
        self.training = True

    def params(self):
        return {k: v for k, v in self.P.items() if v.requires_grad}

    def zero_grad(self):
        for v in self.P.values():
            v.grad = None

    def lower(self, x, ei, batch, G):
        """NodeEmbedding stack + NodeAggregation over one merged chunk."""
        acts = []
        h = x
        for i in range(self.i_agg):
            name, lf = self.specs[i]
            assert name == 'NodeEmbedding'
            h = node_embedding(h, ei, self.P, 'layers.%d' % i, lf, self.training, self.gat_group)
            acts.append(h)
        lf = self.specs[self.i_agg][1]
        multi = _b(lf['concat_multi_scale']) if 'concat_multi_scale' in lf else False
        pooled = readout(acts if multi else [h], batch, G, lf['style'])
        return acts, pooled

    def meta_layer(self, h, etype_eis, p, lf):
        """MetaLayerWrapper + NodeModelAggrByEdge (model/layers_meta.py:39-47,74-79)."""
        kind = lf['node_model'].split('_')[0]
        outs = torch.zeros(h.shape[0], int(lf['output_dim']), dtype=h.dtype, device=h.device)
        for e, ei in enumerate(etype_eis):
            q = p + '.node_model.GNNS.%d' % e
            view = {q + '.conv.' + k: self.P[q + '.' + k] for k in ('weight', 'bias', 'att') if (q + '.' + k) in self.P}
            outs = outs + (gcn_conv(h, ei, view, q) if kind == 'gcn' else gat_conv(h, ei, view, q, self.gat_group))
        return _act(lf['act'], outs)

    def upper(self, init_x, ddi_ei, pair_rows, y, num_labels=2, etype_eis=None):
        h = init_x
        acts = []
        # after LoadInteractionLayer (bi-level), else after the readout (lower-level-only model), else everything
        start = self.i_load + 1 if self.i_load is not None else (self.i_agg + 1 if self.i_agg is not None else 0)
        pred = None
        for i in range(start, len(self.specs)):
            name, lf = self.specs[i]
            p = 'layers.%d' % i
            if name == 'NodeEmbedding':
                h = node_embedding(h, ddi_ei, self.P, p, lf, self.training, self.gat_group)
            elif name == 'MetaLayer':
                h = self.meta_layer(h, etype_eis, p, lf)
            elif name == 'LinkPredictor':
                h = pred = link_pred(h, pair_rows, self.P, p, lf, num_labels)
            elif name == 'Loss':
                h = loss_fn(h, y, lf['type'])
            else:
                raise ValueError('Unknown layer {}'.format(name))
            acts.append(h)
        return acts, pred, h


def all_drug_pass(model, ds, batch_size=64, record=None):
    """src/train.py:48-72 + layers_aggregation.py:70-74: every drug through the
    lower level in chunks; pooled rows scattered into init_x[N, D]."""
    rows_out = [None] * ds.N
    chunks = all_drug_chunks(ds.gids.tolist(), batch_size)
    for c, pairs in enumerate(chunks):
        gids = unique_graphs_in_order(pairs)
        m = merge_graphs(ds, gids)
        dev = getattr(model, 'device', 'cpu')
        x = torch.from_numpy(m['x']).to(model.dtype).to(dev)
        ei = torch.from_numpy(m['edge_index']).to(dev)
        batch = torch.from_numpy(m['batch']).to(dev)
        acts, pooled = model.lower(x, ei, batch, len(gids))
        for g, i in m['gids_to_batch_ind'].items():
            rows_out[ds.gs_map[g]] = pooled[i]
        if record is not None:
            record.append(dict(merge=m, acts=[a.detach() for a in acts], pooled=pooled.detach()))
    return torch.stack(rows_out, 0)


def train_step_forward(model, ds, batch_gids, y, batch_size=64, record=None):
    """model_forward + Model.forward of one Bi-GNN step (src/train.py:75-108,178)."""
    init_x = all_drug_pass(model, ds, batch_size, record)
    dev = getattr(model, 'device', 'cpu')
    ddi = torch.from_numpy(np.stack([ds.ddi_row, ds.ddi_col])).to(dev)
    rows = torch.from_numpy(np.vectorize(ds.gs_map.get)(np.asarray(batch_gids)).astype(np.int64)).to(dev)
    et = [torch.from_numpy(np.stack([r, c])).to(dev) for r, c in ds.etypes.values()] if ds.etypes else None
    acts, pred, loss = model.upper(init_x, ddi, rows, torch.from_numpy(np.asarray(y)).to(dev), etype_eis=et)
    return init_x, acts, pred, loss


def lower_only_step_forward(model, ds, batch_gids, y):
    """The lower-level-only model (model='lower_level_gnn', LL-GNN baseline): src/train.py:99-107 merges
    the pair batch's unique molecule graphs into ONE graph (first-appearance order, src/batch.py:131-136), Model.forward
    runs the NodeEmbedding stack, the readout ([G, L*D], no init_x write: model/layers_aggregation.py:68-69), LinkPred
    over gids_to_batch_ind rows (model/layers_link_pred.py:52) and the loss."""
    gids = unique_graphs_in_order(np.asarray(batch_gids))
    m = merge_graphs(ds, gids)
    dev = getattr(model, 'device', 'cpu')
    x = torch.from_numpy(m['x']).to(model.dtype).to(dev)
    acts, pooled = model.lower(x, torch.from_numpy(m['edge_index']).to(dev), torch.from_numpy(m['batch']).to(dev),
                               len(gids))
    to_row = m['gids_to_batch_ind']
    rows = torch.from_numpy(np.vectorize(to_row.get)(np.asarray(batch_gids)).astype(np.int64)).to(dev)
    _, pred, loss = model.upper(pooled, None, rows, torch.from_numpy(np.asarray(y)).to(dev))
    return m, acts, pooled, pred, loss


def adam_step(P, state, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam defaults (src/train.py:35)."""
    state['t'] = state.get('t', 0) + 1
    t = state['t']
    with torch.no_grad():
        for k, v in P.items():
            if not v.requires_grad or v.grad is None:
                continue
            m = state.setdefault('m/' + k, torch.zeros_like(v))
            s = state.setdefault('v/' + k, torch.zeros_like(v))
            m.mul_(betas[0]).add_(v.grad, alpha=1 - betas[0])
            s.mul_(betas[1]).addcmul_(v.grad, v.grad, value=1 - betas[1])
            bc1 = 1 - betas[0] ** t
            bc2 = 1 - betas[1] ** t
            denom = (s.sqrt() / math.sqrt(bc2)).add_(eps)
            v.addcdiv_(m, denom, value=-lr / bc1)


class OracleTrainer(object):
    """The reference's hot loop body (src/train.py:136-141) over this port, for the CPU
    baseline: per step the all-drug lower pass (per-graph conversion + merge + 5 layers per
    chunk), positive/negative sampling, the upper pass, backward and Adam."""

    def __init__(self, ds, specs, state, batch_size=64, dtype=torch.float32, device='cpu'):
        from torch.utils.data import DataLoader
        self.ds, self.bs = ds, batch_size
        self.model = OracleModel(specs, state, dtype, device=device)
        self.adam = {}
        self.items = torch.as_tensor(np.asarray(sorted(map(tuple, ds.train_pairs.tolist())), np.int64))
        self.loader = DataLoader(self.items, batch_size=batch_size, shuffle=True)
        self.it = iter(self.loader)
        self.edge_set = set(zip(ds.ddi_row.tolist(), ds.ddi_col.tolist()))

    def sample(self):
        try:
            pos = next(self.it)
        except StopIteration:
            self.it = iter(self.loader)
            pos = next(self.it)
        pos = pos.numpy()
        neg = sample_negative_pairs(self.ds, pos, np.unique(pos), self.edge_set)
        gids = np.concatenate([pos, neg]) if len(neg) else pos
        return gids, pair_labels(self.ds, gids)

    def step(self):
        self.model.zero_grad()
        gids, y = self.sample()
        _, _, _, loss = train_step_forward(self.model, self.ds, gids, y, self.bs)
        loss.backward()
        adam_step(self.model.P, self.adam)
        return float(loss.detach()), len(gids)


class OracleLowerOnlyTrainer(OracleTrainer):
    """The lower-level-only model's hot loop (src/train.py:99-107,136-141): sample a pair batch, merge its unique
    molecule graphs graph by graph, GIN stack + readout + scorer, backward, Adam."""

    def step(self):
        self.model.zero_grad()
        gids, y = self.sample()
        _, _, _, _, loss = lower_only_step_forward(self.model, self.ds, gids, y)
        loss.backward()
        adam_step(self.model.P, self.adam)
        return float(loss.detach()), len(gids)


# --------------------------------------------------------------------------
# non-default readouts / heads (model/layers_aggregation.py:45-56,78-95;
# model/layers_util.py:44-57 MLP; model/layers_link_pred.py:66-67)
# --------------------------------------------------------------------------
def mlp_forward(x, P, p, n_layers, bn=False, training=True):
    """model/layers_util.py:44-57: act(bn(linear)) on all but the last layer (relu)."""
    for i in range(n_layers):
        x = F.linear(x, P[p + '.layers.%d.weight' % i], P[p + '.layers.%d.bias' % i])
        if i < n_layers - 1:
            if bn:
                x = F.batch_norm(x, None, None, P[p + '.bn.%d.weight' % i], P[p + '.bn.%d.bias' % i], True, 0.1, 1e-5)
            x = torch.relu(x)
    return x


def deepsets_readout(x, batch, G, P, p, n_layers):
    h = mlp_forward(x, P, p + '.phi', n_layers)
    h = readout([h], batch, G, 'avg_pool')
    return mlp_forward(h, P, p + '.rho', n_layers)


def gmn_aggr_readout(x, batch, G, P, p):
    w = mlp_forward(x, P, p + '.weight_func', 2, bn=True)
    g = torch.sigmoid(mlp_forward(x, P, p + '.gate_func', 2, bn=True))
    emb = _scatter_rows(g * w, batch, G)
    return mlp_forward(emb, P, p + '.mlp_graph', 2, bn=True)
