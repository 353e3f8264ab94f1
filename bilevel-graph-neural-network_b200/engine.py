"""BiGNNEngine: one Bi-GNN train step as a single batched, CUDA-graph-captured pass.

Same arithmetic as the reference-sequenced driver in train.py (and therefore as
src/train.py:75-108,177-182), re-organised for the GPU:

  * the all-drug lower pass (src/train.py:48-72: 11 chunks x 5 layers on DrugBank) runs as ONE
    merged graph over every drug; BatchNorm keeps the reference's per-chunk statistics through
    the segmented kernels (`chunk_row_ptr`), so results equal the chunk-by-chunk loop;
  * the readout writes the pooled rows straight into init_x[gs_map[gid]] (dst_row) instead of
    128 Python row copies per chunk (model/layers_aggregation.py:70-74);
  * per-step inputs (pair rows, labels, decoder entry CSR) live in static device buffers fed
    from pinned host memory, and forward + backward + Adam are replayed as one CUDA graph.
"""
import numpy as np
import torch

import os

from . import ops, dist as bdist, fused
from .batch import BatchData, unique_graphs_in_order
from .config import get_flags
from .graph import MergedGraph, PartitionedInteractionGraph
from .layers_load_interaction_graph import LoadInteractionGraph
from .ops import CSR
from .train import all_drug_chunks


class _StaticPairBatch(object):
    """The pair batch as the upper-level layers see it (LinkPred / Loss), backed by static
    device buffers so that a captured graph can be replayed on new pairs."""

    def __init__(self, data, P, device, graph=None):
        self.dataset = data
        self.P = P
        # pair rows index the matrix the scorer gathers from: the N drug rows, or -- with a row-partitioned
        # upper level -- the world*n_max positions of the all-gathered embeddings
        N = data.N if graph is None else graph.n_pad
        self.interaction_combo_nxgraph = data.interaction_combo_nxgraph if graph is None else graph
        self.merge_data = {}
        self.merge_higher_level = {}
        self.ids = torch.zeros((P, 2), dtype=torch.int32, device=device)
        self.y = torch.zeros(P, dtype=torch.float32, device=device)
        # decoder entry CSR (drug row -> the pair entries that gather it): the 2P sorted rows come from the host,
        # the N+1 pointer array is derived from them ON THE DEVICE each step (`refresh`), not copied
        self.e_rows = torch.zeros(2 * P, dtype=torch.int32, device=device)
        self.e_ptr = torch.zeros(N + 1, dtype=torch.int32, device=device)
        self.e_idx = torch.zeros(2 * P, dtype=torch.int32, device=device)
        self._bounds = torch.arange(N + 1, dtype=torch.int32, device=device)
        self.entry_csr = CSR(self.e_ptr, self.e_idx, N)
        self.chk = torch.zeros(2, dtype=torch.float32, device=device)      # pair-batch checksum (multi-GPU agreement)
        self.batch_gids = np.zeros((P, 2), np.int64)
        self.preds = None

    def load(self, st, refresh=True):
        """staged inputs (pinned host) -> the static device buffers."""
        self.ids.copy_(st.ids, non_blocking=True)
        self.y.copy_(st.y, non_blocking=True)
        self.e_rows.copy_(st.e_rows, non_blocking=True)
        self.e_idx.copy_(st.e_idx, non_blocking=True)
        self.chk.copy_(st.chk, non_blocking=True)
        if refresh:
            self.refresh()

    def refresh(self):
        """e_ptr[r] = number of entries whose row is < r (capturable; runs at the head of every step)."""
        torch.searchsorted(self.e_rows, self._bounds, out_int32=True, out=self.e_ptr)

    def y_true_device(self, as_int=False):
        return self.y.to(torch.int32) if as_int else self.y

    def pair_rows_device(self, n_rows, higher=True, unique=True):
        return self.ids, self.entry_csr

    def assign_link_preds(self, pair_preds):
        self.preds = pair_preds.detach()


class _Staging(object):
    """Pinned host buffers for one step's inputs (+ the loss read-back)."""

    def __init__(self, P, N, cuda=True):
        pin = dict(pin_memory=True) if cuda else {}
        self.ids = torch.zeros((P, 2), dtype=torch.int32, **pin)
        self.y = torch.zeros(P, dtype=torch.float32, **pin)
        self.e_rows = torch.zeros(2 * P, dtype=torch.int32, **pin)
        self.e_idx = torch.zeros(2 * P, dtype=torch.int32, **pin)
        self.chk = torch.zeros(2, dtype=torch.float32, **pin)
        self.loss = torch.zeros((), dtype=torch.float32, **pin)
        self.chk_sum = torch.zeros(2, dtype=torch.float32, **pin)
        self.event = torch.cuda.Event() if cuda else None
        self.busy = False

    def nbytes_in(self):
        return sum(t.numel() * t.element_size() for t in (self.ids, self.y, self.e_rows, self.e_idx, self.chk))


class BiGNNEngine(object):
    def __init__(self, data, model, optimizer=None, lr=None, use_cuda_graph=True, rebuild_each_step=True,
                 n_staging=4, rank=0, world=1, group=None, adam_capturable=None, partition_upper=None,
                 fused_lower=None):
        flags = get_flags()
        assert flags.lower_level_layers and flags.higher_level_layers, 'engine runs the Bi-GNN mode'
        self.data, self.model = data, model
        self.device = data.device
        self.use_cuda_graph = use_cuda_graph and self.device.type == 'cuda'
        self.rebuild_each_step = rebuild_each_step
        self.optimizer = optimizer if optimizer is not None else torch.optim.Adam(
            model.parameters(), lr=flags.lr if lr is None else lr,
            capturable=self.use_cuda_graph if adam_capturable is None else adam_capturable)
        # ---- static all-drug merged graph: chunk schedule of src/train.py:52-71
        self.rank, self.world, self.group = int(rank), int(world), group
        gids = list(data.gs_map.keys())
        chunk_rows = []
        for pairs in all_drug_chunks(gids, flags.batch_size):
            order = unique_graphs_in_order(pairs)
            chunk_rows.append(np.asarray([data.gs_map[g] for g in order.keys()], np.int64))
        self.n_chunks_total = len(chunk_rows)
        # shard whole chunks over the ranks, balanced by atoms (dist.shard_chunks)
        weights = [int(data.packed.sizes(r)[0].sum()) for r in chunk_rows]
        self.chunk_shards = bdist.shard_chunks(weights, self.world)
        lo, hi = self.chunk_shards[self.rank]
        self.my_chunks = (lo, hi)
        all_rows_in_order = np.concatenate(chunk_rows)
        mine = chunk_rows[lo:hi]
        rows = np.concatenate(mine) if mine else np.zeros(0, np.int64)
        chunk_ptr = np.concatenate([[0], np.cumsum([len(r) for r in mine])]).astype(np.int64)
        # the lower level runs as fused layer kernels (fused.py) whenever they cover it exactly; otherwise -- and with
        # BIGNN_NO_FUSED=1 -- layer by layer through the registry's layer classes
        agg = model.lower_layers[-1]
        self.fused_lower = (self.device.type == 'cuda' and not os.environ.get('BIGNN_NO_FUSED')
                            and fused.stack_supported(model.init_layers, agg, data.num_node_feat)) \
            if fused_lower is None else bool(fused_lower)
        self.lower_path = 'fused' if self.fused_lower else 'layers'
        self.merged = MergedGraph(data.packed, rows, chunk_graph_ptr=chunk_ptr, pad_features=self.fused_lower)
        self._bn_sink = [] if self.world > 1 else None
        self.merged.bn_stats_sink = self._bn_sink
        # ---- upper level: replicated, or rows (= edges by source drug) partitioned over the ranks with a
        # per-layer all-gather.  Partitioned by default when every upper layer is a GCN.
        from .layers import NodeEmbedding
        convs = [l for l in model.higher_level_layers if not isinstance(l, (LoadInteractionGraph,))][:-2]
        can_part = all(isinstance(l, NodeEmbedding) and l.type == 'gcn' for l in convs)
        if partition_upper is None:
            partition_upper = self.world > 1 and can_part
        if partition_upper and not can_part:
            raise NotImplementedError('partition_upper: only GCN upper levels are row-partitioned')
        self.upper = PartitionedInteractionGraph(data.interaction_combo_nxgraph, self.rank, self.world,
                                                 group) if partition_upper else None
        # parameters evaluated on this rank's rows only hold partial gradient sums: the conv weights / biases and a
        # parametric activation (PReLU) join the flat all-reduce; BatchNorm's come out rank-summed already
        self._upper_partial = [p for l in convs for m in (l.conv, l.act) for p in m.parameters()] \
            if partition_upper else []
        self._n_pair_rows = self.upper.n_pad if self.upper is not None else data.N
        # a drug that appears in two chunks is overwritten by the later one in the reference
        # (layers_aggregation.py:72-74); earlier duplicates go to a trash row
        first_later = {}
        for i in range(len(all_rows_in_order) - 1, -1, -1):
            first_later.setdefault(int(all_rows_in_order[i]), i)      # last occurrence wins
        offset = int(sum(len(r) for r in chunk_rows[:lo]))
        if self.world == 1:
            dst = rows.copy()
            for i in range(len(rows)):
                if first_later[int(rows[i])] != offset + i:
                    dst[i] = data.N
            self._pool_rows = data.N + 1
            self.pooled_layout = None
        else:
            # exchange step = ALL-GATHER of pooled drug rows: every rank pools into its own compact block
            # [n_max (+1 trash), L*D] (drug rows it owns, ascending); block r of the gathered [world, n_max, L*D] holds
            # rank r's drugs, and `perm` maps drug row -> position in it
            owned, off = [], 0
            for (clo, chi) in self.chunk_shards:
                cnt = int(sum(len(r) for r in chunk_rows[clo:chi]))
                seg = all_rows_in_order[off:off + cnt]
                keep = np.asarray([first_later[int(g)] == off + i for i, g in enumerate(seg)], bool) \
                    if cnt else np.zeros(0, bool)
                owned.append(np.sort(seg[keep]))
                off += cnt
            n_max = max(max(len(o) for o in owned), 1)
            perm = np.zeros(data.N, np.int64)
            for r, o in enumerate(owned):
                perm[o] = r * n_max + np.arange(len(o))
            mine_sorted = owned[self.rank]
            dst = np.full(len(rows), n_max, np.int64)                  # trash row of the local block
            for i in range(len(rows)):
                if first_later[int(rows[i])] == offset + i:
                    dst[i] = int(np.searchsorted(mine_sorted, rows[i]))
            self._pool_rows = n_max + 1
            self.pooled_layout = bdist.PooledLayout(self.rank, self.world, n_max, len(mine_sorted),
                                                    torch.as_tensor(perm).to(self.device), group)
        self.dst_row = torch.as_tensor(dst.astype(np.int32)).to(self.device)
        self._all_chunk_ptr = torch.arange(self.n_chunks_total + 1, dtype=torch.int32, device=self.device)
        self._lower_bd = type('LowerBatch', (), {})()
        self._lower_bd.merge_data = {'merge': self.merged}
        self._lower_bd.merge_higher_level = {}
        self._lower_bd.dataset = data
        self._agg_style, self._multi = agg.style, agg.concat_multi_scale
        self._stack = fused.StackSpec(list(model.init_layers), agg, self.merged, self.dst_row, self._pool_rows,
                                      self._bn_sink) if self.fused_lower else None
        self._graphs = {}          # P -> (graph, static batch, loss tensor)
        self.max_graphs = 8
        self._stagings = {}
        self._n_staging = n_staging
        self._step_idx = 0
        self._chk_sum = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.init_x_static = None      # the last step's init_x (detached copy; what validation scores, src/train.py:185-220)
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 4

    # ------------------------------------------------------------------ device work
    def lower_pass(self):
        """All drugs through the lower level -> init_x [N(+1), L*D] (with autograd history)."""
        m, model = self.merged, self.model
        if self.rebuild_each_step:
            m.build()
        acts = []
        if self._stack is not None:
            pooled = fused.gin_stack(self._stack, m.x, model.training)
        else:
            h = m.x
            for layer in model.init_layers:
                h = layer(h, self._lower_bd, model)
                acts.append(h)
            pooled = ops.readout(acts if self._multi else [h], m.seg_ptr, m.G, self._agg_style,
                                 self.dst_row, self._pool_rows)
        if self.world > 1:                                            # the exchange step (NCCL all-gather)
            pooled = bdist.gather_pooled_rows(pooled, self.pooled_layout, partial_grad=self.upper is not None)
        return pooled, acts

    def _ig(self):
        return self.upper if self.upper is not None else self.data.interaction_combo_nxgraph

    def forward(self, pair_batch):
        model = self.model
        pooled, _ = self.lower_pass()
        self._ig().init_x = pooled[:self.data.N]
        with torch.no_grad():                                # engine-owned copy: survives CUDA-graph pool reuse
            if self.init_x_static is None:
                self.init_x_static = torch.empty_like(pooled[:self.data.N])
            self.init_x_static.copy_(pooled[:self.data.N])
        model.use_layers = 'higher_layers'
        model.acts = [None]
        for layer in model.higher_level_layers:
            model.acts.append(layer(model.acts[-1], pair_batch, model))
        return model.acts[-1]

    def _sync_lower(self, pair_batch=None):
        """after backward on a sharded lower level: sum the partial weight gradients and replay
        the BatchNorm running-buffer updates in global chunk order."""
        lower = [p for l in self.model.init_layers for p in l.parameters()]
        # + the conv weights / biases of a row-partitioned upper level (partial sums over a rank's rows)
        # + the pair-batch checksum [c, c^2]: every rank must have staged the SAME pairs (the scorer and the upper
        #   level treat their gradients as replicated); world * sum(c^2) == (sum c)^2 iff all ranks agree
        extra = pair_batch.chk if pair_batch is not None and hasattr(pair_batch, 'chk') else None
        red = bdist.all_reduce_grads(lower + self._upper_partial, self.group, extra=extra)
        if red is not None:
            self._chk_sum.copy_(red)
        s_max = max(hi - lo for lo, hi in self.chunk_shards)
        if self._bn_sink:
            # the per-chunk statistics of ALL BatchNorm layers travel in one all-gather
            widths = [st.shape[2] for _, st in self._bn_sink]
            if len(set(widths)) == 1:
                stacked = torch.cat([st for _, st in self._bn_sink], dim=0)              # [2L, S_local, C]
                allst = bdist.gather_chunk_stats(stacked, s_max, self.group)             # [world, 2L, s_max, C]
                per_layer = [allst[:, 2 * l:2 * l + 2] for l in range(len(self._bn_sink))]
            else:
                per_layer = [bdist.gather_chunk_stats(st, s_max, self.group) for _, st in self._bn_sink]
            for (bn, _), allst in zip(self._bn_sink, per_layer):
                parts = [allst[r][:, :hi - lo] for r, (lo, hi) in enumerate(self.chunk_shards)]
                ordered = torch.cat(parts, dim=1).contiguous()                           # [2, S_total, C]
                ops.bn_running_update(ordered, self._all_chunk_ptr, self.n_chunks_total, bn.running_mean,
                                      bn.running_var, bn.num_batches_tracked, bn.momentum)
        del self._bn_sink[:]

    def _device_step(self, pair_batch):
        self.optimizer.zero_grad(set_to_none=True)
        pair_batch.refresh()
        loss = self.forward(pair_batch)
        loss.backward()
        if self.world > 1:
            self._sync_lower(pair_batch)
        self.optimizer.step()
        self._detach_init_x()
        return loss.detach()

    def _detach_init_x(self):
        ig = self._ig()
        if self.upper is not None:
            ig.init_x_full = ig.init_x_full.detach()
        else:
            ig.init_x = ig.init_x.detach()

    # ------------------------------------------------------------------ capture
    def _snapshot(self):
        snap = {'model': {k: v.detach().clone() for k, v in self.model.state_dict().items()}}
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(snap['model'][k])
            for st in self.optimizer.state.values():
                for v in st.values():
                    if isinstance(v, torch.Tensor):
                        v.zero_()

    def _capture(self, P):
        sb = _StaticPairBatch(self.data, P, self.device, self.upper)
        # a valid dummy batch for warm-up: pairs (0,1) with label 0
        sb.ids[:, 1] = 1
        sb.e_rows[P:] = 1
        sb.e_idx.copy_(torch.as_tensor(np.concatenate([np.arange(0, 2 * P, 2), np.arange(1, 2 * P, 2)]).astype(np.int32)))
        had_state = len(self.optimizer.state) > 0
        snap = self._snapshot()
        opt_snap = None
        if had_state:
            opt_snap = [{k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
                        for st in self.optimizer.state.values()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._device_step(sb)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        from . import _lib
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            loss = self._device_step(sb)
        self.launches_per_step = _lib.launch_count() - n0
        # undo the warm-up steps: parameters, BN buffers and Adam moments back to where they were
        self._restore(snap)
        if opt_snap is not None:
            with torch.no_grad():
                for st, old in zip(self.optimizer.state.values(), opt_snap):
                    for k, v in st.items():
                        if isinstance(v, torch.Tensor):
                            v.copy_(old[k])
        torch.cuda.synchronize()
        self._graphs[P] = (g, sb, loss)
        return self._graphs[P]

    # ------------------------------------------------------------------ host side of a step
    def _staging(self, P):
        key = P
        if key not in self._stagings:
            if len(self._stagings) >= 16:                  # varying batch sizes: keep the pinned pool bounded
                old = next(iter(self._stagings))
                for s_ in self._stagings.pop(old):
                    if s_.busy and s_.event is not None:
                        s_.event.synchronize()
            self._stagings[key] = [_Staging(P, self._n_pair_rows, self.device.type == 'cuda')
                                   for _ in range(self._n_staging)]
        st = self._stagings[key][self._step_idx % self._n_staging]
        if st.busy and st.event is not None:
            st.event.synchronize()
        st.busy = False
        return st

    def _pair_rows(self, batch_gids):
        """gid pairs -> flat rows of the matrix the scorer gathers from (drug rows; positions of the
        all-gathered embeddings when the upper level is row-partitioned)."""
        gs_map = self.data.gs_map
        flat = np.fromiter((gs_map[g] for g in np.asarray(batch_gids).reshape(-1).tolist()), np.int64)
        return flat if self.upper is None else self.upper.part.pos(flat)

    def stage_pairs(self, batch_gids, labels):
        """Host -> pinned staging: pair rows (gs_map), labels and the decoder's entry CSR."""
        gs_map = self.data.gs_map
        flat = self._pair_rows(batch_gids)
        P = flat.shape[0] // 2
        st = self._staging(P)
        st.ids.numpy()[:] = flat.reshape(P, 2)
        st.y.numpy()[:] = labels
        order = np.argsort(flat, kind='stable')
        st.e_rows.numpy()[:] = flat[order]
        st.e_idx.numpy()[:] = order
        c = float((int(np.dot(flat % 1021, np.arange(1, flat.shape[0] + 1) % 1019)) +
                   int(np.asarray(labels, np.int64).sum())) % 1009)
        st.chk.numpy()[:] = (c, c * c)
        return st, P

    def step_staged(self, st, P):
        """Runs one train step on staged inputs; returns the staging slot whose `.loss`
        holds the loss once `.event` has completed (read it with `read_loss`)."""
        # one captured graph per pair-batch size; samplers with a varying batch (NeighborSampler, the short last
        # batch of an epoch) get at most `max_graphs` of them, further sizes run as eager launches
        use_graph = self.use_cuda_graph and (
            P in self._graphs or sum(1 for k in self._graphs if isinstance(k, int)) < self.max_graphs)
        if use_graph:
            entry = self._graphs.get(P) or self._capture(P)
            g, sb, loss = entry
        else:
            sb = self._graphs.get(('eager', P))
            if sb is None:
                for k in [k for k in self._graphs if isinstance(k, tuple)]:
                    del self._graphs[k]                    # keep one eager batch
                sb = self._graphs[('eager', P)] = _StaticPairBatch(self.data, P, self.device, self.upper)
        sb.load(st, refresh=False)               # (the pointer array is derived inside the step)
        if use_graph:
            g.replay()
        else:
            loss = self._device_step(sb)
        st.loss.copy_(loss, non_blocking=True)
        if self.world > 1:
            st.chk_sum.copy_(self._chk_sum, non_blocking=True)
        st.world = self.world
        if st.event is not None:
            st.event.record()
        st.busy = True
        self._step_idx += 1
        self.h2d_bytes_per_step = st.nbytes_in()
        self.last_static_batch = sb
        return st

    @staticmethod
    def read_loss(st):
        if st.event is not None:
            st.event.synchronize()
        st.busy = False
        w = getattr(st, 'world', 1)
        if w > 1:
            a, b = float(st.chk_sum[0]), float(st.chk_sum[1])
            if abs(w * b - a * a) > 0.5:
                raise RuntimeError('BiGNNEngine: the ranks staged different pair batches in this step (every rank must '
                                   'sample the same pairs: seed numpy and torch identically on all ranks, see '
                                   'INTEGRATION.md)')
        return float(st.loss)

    @torch.no_grad()
    def score_pairs(self, gid_pairs, recompute_init_x=True):
        """Evaluation path (src/train.py:185-220): model.eval() statistics, one all-drug lower pass
        (or the last training init_x, as the reference's validation does), one upper pass over the whole
        interaction graph and ONE decoder launch over all given pairs -- instead of a full upper pass and
        128 `.item()` syncs per 64-pair batch.  Returns the LinkPred outputs [P, 1] (or [P, K] logits)."""
        model = self.model
        was_training = model.training
        model.eval()
        try:
            ig = self._ig()
            if recompute_init_x or (self.init_x_static is None and ig.init_x is None):
                pooled, _ = self.lower_pass()                         # eval-mode BatchNorm (running statistics)
                ig.init_x = pooled[:self.data.N]
            elif self.init_x_static is not None:
                ig.init_x = self.init_x_static                        # the last TRAIN step's init_x, as the
                                                                      # reference's validation uses (train.py:185-220)
            flat = self._pair_rows(gid_pairs)
            P = flat.shape[0] // 2
            sb = _StaticPairBatch(self.data, P, self.device, self.upper)
            sb.ids.copy_(torch.as_tensor(flat.reshape(P, 2).astype(np.int32)))
            order = np.argsort(flat, kind='stable')
            sb.e_rows.copy_(torch.as_tensor(flat[order].astype(np.int32)))
            sb.e_idx.copy_(torch.as_tensor(order.astype(np.int32)))
            sb.refresh()
            model.acts = [None]
            for layer in model.higher_level_layers[:-1]:           # everything but the loss
                model.acts.append(layer(model.acts[-1], sb, model))
            return model.acts[-1]
        finally:
            model.train(was_training)

    def train_step(self, sampler):
        """Public API: sample (host, bit-exact with the reference), stage, run, return the
        staging slot (loss is read back asynchronously)."""
        batch_gids, sampled_gids, _ = sampler.sample_next_training_batch()
        bd = BatchData(batch_gids, self.data, sampled_gids=sampled_gids, is_train=True, merge_graphs=False)
        labels = np.asarray([p.true_label for p in bd.pair_list], np.float32)
        st, P = self.stage_pairs(bd.batch_gids, labels)
        self.last_batch = bd
        return self.step_staged(st, P)
