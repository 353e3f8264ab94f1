"""LowerOnlyEngine: the lower-level-only model (model='lower_level_gnn', the LL-GNN baseline; BASELINE config 3) as
one batched train step.

The reference's step (src/train.py:99-107 with `lower_level_layers` only) merges the unique molecule graphs of the
pair batch into ONE graph (first-appearance order, src/batch.py:131-136), runs the GIN stack and the readout over it
(one BatchNorm batch), scores the pairs through `gids_to_batch_ind` (model/layers_link_pred.py:52) and steps Adam.
This engine does exactly that, for pair batches of any size: the merged graph is built on the device from the packed
dataset (`bignn_merge_build`), the stack runs as the fused layer kernels (fused.py) when they cover it, the readout
writes [G, L*D], and the scorer / loss / backward / Adam follow.  Batch sizes vary from step to step, so the step is a
sequence of eager launches (each of them large at the batch sizes this engine is for).

`fast_negative_pairs` is a vectorised negative sampler for pair batches of tens of thousands (the reference's
rejection loop is a Python loop with one `np.random.choice` per attempt, src/batch.py:61-103: 0.6 s of host time for
32 768 positives).  It applies the same rejection rules (no self pairs, no duplicates in either orientation, no known
interaction in either orientation, candidates among the batch's own drugs) but draws from its own generator: it is NOT
bit-compatible with the reference's sample stream -- `batch.sample_negative_pairs` is, and stays the default.
"""
import numpy as np
import torch

from . import ops, fused
from .batch import sample_negative_pairs
from .config import get_flags
from .graph import MergedGraph, entry_csr


def fast_negative_pairs(data, positive_gids, rng):
    """One negative (orig, cand) per positive, both among the batch's drugs; see the module docstring."""
    pos = np.asarray(positive_gids, np.int64)
    sampled = np.unique(pos)
    n, P = sampled.shape[0], pos.shape[0]
    rows_of = data.rows_of_gids
    N = data.N
    s_rows = rows_of(sampled)
    pos_rows = rows_of(pos.reshape(-1)).reshape(-1, 2)
    pos_keys = np.unique(np.concatenate([pos_rows[:, 0] * N + pos_rows[:, 1], pos_rows[:, 1] * N + pos_rows[:, 0]]))
    edge_keys = data.edge_keys_sorted()

    def known(a, b):
        k = a * N + b
        i = np.searchsorted(edge_keys, k)
        hit = (i < edge_keys.shape[0]) & (edge_keys[np.minimum(i, edge_keys.shape[0] - 1)] == k)
        j = np.searchsorted(pos_keys, k)
        return hit | ((j < pos_keys.shape[0]) & (pos_keys[np.minimum(j, pos_keys.shape[0] - 1)] == k))
    target = int(min(n * (n - 1) // 2 - P, P))
    out_a, out_b, seen = [], [], np.zeros(0, np.int64)
    need, cursor = target, 0
    while need > 0:
        m = int(need * 1.2) + 16
        a_i = (cursor + np.arange(m)) % n                  # the reference walks `orig` cyclically over the batch's drugs
        cursor += m
        b_i = rng.integers(0, n, m)
        a, b = s_rows[a_i], s_rows[b_i]
        ok = (a != b) & ~known(a, b) & ~known(b, a)
        key = np.minimum(a, b) * N + np.maximum(a, b)      # orientation-free: (a, b) and (b, a) are the same negative
        ok &= ~np.isin(key, seen)
        _, first = np.unique(key, return_index=True)
        uniq = np.zeros(m, bool)
        uniq[first] = True
        ok &= uniq
        idx = np.nonzero(ok)[0][:need]
        out_a.append(sampled[a_i[idx]])
        out_b.append(sampled[b_i[idx]])
        seen = np.concatenate([seen, key[idx]])
        need -= idx.shape[0]
    if not out_a:
        return np.zeros((0, 2), np.int64)
    return np.stack([np.concatenate(out_a), np.concatenate(out_b)], 1)


class FastPairSampler(object):
    """Shuffled mini-batches of positive train pairs for LARGE batches: an epoch permutation drawn with one
    `torch.randperm` and sliced, instead of a `DataLoader` that collates 32 768 two-element tensors one by one (100 ms of
    host time per batch).  Same distribution as src/sampler.py:110-131, not the same random stream."""

    def __init__(self, data, batch_size, seed=0):
        self.pairs = np.asarray(data.train_pairs, np.int64)
        self.batch_size = int(batch_size)
        self.gen = torch.Generator().manual_seed(int(seed))
        self._perm, self._at = None, 0

    def sample_next_training_batch(self):
        n = self.pairs.shape[0]
        if self._perm is None or self._at >= n:
            self._perm, self._at = torch.randperm(n, generator=self.gen).numpy(), 0
        idx = self._perm[self._at:self._at + self.batch_size]
        self._at += self.batch_size
        pos = self.pairs[idx]
        return pos, None, None                      # (the engine derives the batch's drugs itself)


class PairPrefetcher(object):
    """Prepares the pair batches of the coming steps on host threads while the device runs.

    At 32 768 + 32 768 pairs per step the host side of a step (vectorised negatives 65 ms, labels 49 ms, unique drugs in
    first-appearance order 43 ms -- all of it binary searches and sorts that release the GIL) costs seven times the
    device step (22 ms on B200).  Positives are taken from `sampler` in order on the calling thread; batch i's negatives
    come from a generator seeded with (seed, i), so the stream of batches does not depend on the number of workers.
    `next()` returns ((rows, ids, labels), number of pairs) of the next step, in order."""

    def __init__(self, engine, sampler, workers=None, depth=None, seed=0):
        import collections
        import os
        from concurrent.futures import ThreadPoolExecutor
        self.eng, self.sampler, self.seed = engine, sampler, int(seed)
        self.workers = int(workers) if workers else max(1, min(16, (os.cpu_count() or 2) - 1))
        self.depth = int(depth) if depth else 2 * self.workers
        self.pool = ThreadPoolExecutor(self.workers)
        self.q = collections.deque()
        self.n = 0
        d = engine.data                     # build the lazily cached lookup tables before any worker needs them
        d.rows_of_gids(np.zeros(0, np.int64))
        d.edge_keys_sorted()

    def prepare(self, pos, i):
        d = self.eng.data
        neg = fast_negative_pairs(d, pos, np.random.default_rng([self.seed, int(i)]))
        gids = np.concatenate([pos, neg]) if neg.shape[0] else pos
        labels = np.concatenate([d.labels_of_pairs(pos), np.zeros(neg.shape[0], np.int64)])
        return self.eng.stage(gids, labels), int(gids.shape[0])

    def _fill(self):
        while len(self.q) < self.depth:
            pos = np.asarray(self.sampler.sample_next_training_batch()[0], np.int64)
            self.q.append(self.pool.submit(self.prepare, pos, self.n))
            self.n += 1

    def next(self):
        self._fill()
        out = self.q.popleft().result()
        self._fill()
        return out

    def close(self):
        self.pool.shutdown(wait=False, cancel_futures=True)
        self.q.clear()


class _LowerPairBatch(object):
    """What LinkPred / Loss read of a batch in the lower-level-only model."""

    def __init__(self, data, merged, ids, labels, device):
        self.dataset = data
        self.merge_data = {'merge': merged}
        self.merge_higher_level = {}
        self.interaction_combo_nxgraph = data.interaction_combo_nxgraph
        self._ids = torch.as_tensor(ids.astype(np.int32)).to(device, non_blocking=True)
        self._csr = entry_csr(ids, merged.G, device)
        y = np.asarray(labels)
        self._y = (torch.as_tensor(y.astype(np.float32)).to(device, non_blocking=True),
                   torch.as_tensor(y.astype(np.int32)).to(device, non_blocking=True))
        self.preds = None

    def y_true_device(self, as_int=False):
        return self._y[1] if as_int else self._y[0]

    def pair_rows_device(self, n_rows, higher=True, unique=True):
        return self._ids, self._csr

    def assign_link_preds(self, pair_preds):
        self.preds = pair_preds.detach()


class LowerOnlyEngine(object):
    def __init__(self, data, model, optimizer=None, lr=None, fused_lower=None):
        flags = get_flags()
        assert flags.lower_level_layers and not flags.higher_level_layers, 'engine runs the lower-level-only model'
        self.data, self.model, self.device = data, model, data.device
        self.optimizer = optimizer if optimizer is not None else torch.optim.Adam(
            model.parameters(), lr=flags.lr if lr is None else lr)
        self.gin_layers = list(model.init_layers[:-1])
        self.agg = model.init_layers[-1]
        self.scorer, self.loss_layer = model.layers[-2], model.layers[-1]
        import os
        self.fused_lower = (self.device.type == 'cuda' and not os.environ.get('BIGNN_NO_FUSED')
                            and fused.stack_supported(self.gin_layers, self.agg, data.num_node_feat)) \
            if fused_lower is None else bool(fused_lower)
        self.lower_path = 'fused' if self.fused_lower else 'layers'
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 4
        self.last = None

    # ------------------------------------------------------------------ host side
    def stage(self, batch_gids, labels):
        """gid pairs -> (rows of the unique drugs in first-appearance order, [P, 2] positions in that order)."""
        flat = self.data.rows_of_gids(np.asarray(batch_gids, np.int64).reshape(-1))
        uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
        order = np.argsort(first, kind='stable')
        rank = np.empty_like(order)
        rank[order] = np.arange(order.shape[0])
        ids = rank[inv].reshape(-1, 2)
        return uniq[order], ids, np.asarray(labels)

    # ------------------------------------------------------------------ device side
    def forward(self, rows, ids, labels):
        model = self.model
        m = MergedGraph(self.data.packed, rows, pad_features=self.fused_lower)
        pb = _LowerPairBatch(self.data, m, ids, labels, self.device)
        self.h2d_bytes_per_step = int(rows.shape[0] * 4 + ids.size * 4 + 2 * ids.size * 4 + (m.G + 1) * 4 + labels.shape[0] * 8)
        if self.fused_lower:
            spec = fused.StackSpec(self.gin_layers, self.agg, m, None, m.G, None)
            pooled = fused.gin_stack(spec, m.x, model.training)
        else:
            acts, h = [], m.x
            for layer in self.gin_layers:
                h = layer(h, pb, model)
                acts.append(h)
            pooled = ops.readout(acts if self.agg.concat_multi_scale else [h], m.seg_ptr, m.G, self.agg.style)
        scores = self.scorer(pooled, pb, model)
        loss = self.loss_layer(scores, pb, model)
        self.last = dict(merged=m, pooled=pooled, scores=scores, batch=pb)
        return loss

    use_cuda_graph = False          # batch sizes change from step to step: eager launches

    def _device_step(self, staged):
        """forward + backward + Adam on a staged batch (rows, ids, labels)."""
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.forward(*staged)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def step_pairs(self, batch_gids, labels):
        self.last_static_batch = self.stage(batch_gids, labels)
        return self._device_step(self.last_static_batch)

    def train_step_prefetched(self, prefetcher):
        """one step on the next batch of a PairPrefetcher (host preparation of later batches overlaps this step)."""
        staged, n = prefetcher.next()
        self.last_pairs = n
        self.last_static_batch = staged
        return self._device_step(staged)

    def train_step(self, sampler, fast_negatives=None, rng=None):
        """sample positives (the reference's DataLoader mechanism), negatives, labels; run one step.
        fast_negatives: vectorised sampler (default for batches above 4 096 positives; not bit-compatible)."""
        pos, sampled, _ = sampler.sample_next_training_batch()
        pos = np.asarray(pos, np.int64)
        if fast_negatives is None:
            fast_negatives = pos.shape[0] > 4096
        if fast_negatives:
            neg = fast_negative_pairs(self.data, pos, rng if rng is not None else np.random.default_rng(pos.shape[0]))
            gids = np.concatenate([pos, neg]) if neg.shape[0] else pos
            labels = np.concatenate([self.data.labels_of_pairs(pos), np.zeros(neg.shape[0], np.int64)])
        else:
            flags = get_flags()
            neg = sample_negative_pairs(self.data, pos, sampled, flags.num_negative_samples)
            gids = np.concatenate([pos, neg]) if len(neg) else pos
            labels = np.asarray([0 if l is None else l for l in
                                 (self.data.look_up_label(int(a), int(b)) for a, b in gids.tolist())], np.int64)
        self.last_pairs = gids.shape[0]
        return self.step_pairs(gids, labels)
