"""BatchData (src/batch.py:14-162): one pair mini-batch -- negative sampling, unique
graphs in first-appearance order, pair labels, and the merged molecule graph, which is
built on the device (bignn_merge_build) instead of per graph on the host."""
import numpy as np
import torch

from .config import get_flags
from .graph import MergedGraph, entry_csr


def unique_graphs_in_order(batch_gids):
    """dict-insertion order over pairs scanned row-major (src/batch.py:112-113,131-136)."""
    seen = {}
    for g1, g2 in np.asarray(batch_gids).tolist():
        if g1 not in seen:
            seen[g1] = len(seen)
        if g2 not in seen:
            seen[g2] = len(seen)
    return seen


def sample_negative_pairs(dataset, positive_gids, sampled_gids, num_negative_samples=1, rng=np.random,
                          enforce_negative=True, amongst_same_graphs=True):
    """src/batch.py:61-103, verbatim in its decisions: target count, cyclic `orig` walk, one
    `np.random.choice(negative_gids, size=1)` per attempt, the five rejection tests, and the
    CPython set -> dict -> list result order.  Only the membership structure is hoisted:
    the train-edge set is built once per dataset, not once per call.
    enforce_negative=False skips the rejection loop (src/batch.py:88); amongst_same_graphs=False
    (FLAGS.enforce_sampling_amongst_same_graphs, :69-79) draws candidates from every drug and lifts the pair cap."""
    neg = set()
    gid_ind = 0
    n = len(sampled_gids)
    gs_map = dataset.gs_map
    max_pairs = ((n * (n - 1)) / 2) - len(positive_gids) if amongst_same_graphs else float('inf')
    negative_gids = sampled_gids if amongst_same_graphs else list(gs_map.keys())
    edges = dataset.edge_set()
    batch_pos = set((gs_map[int(a)], gs_map[int(b)]) for a, b in positive_gids)
    target = min(max_pairs, len(positive_gids) * num_negative_samples)

    def is_pos(a, b):
        k = (gs_map[int(a)], gs_map[int(b)])
        return k in edges or k in batch_pos

    while True:
        if len(neg) == target:
            break
        orig = sampled_gids[gid_ind % n]
        cand = rng.choice(negative_gids, size=1)[0]
        gid_ind += 1
        while enforce_negative and ((orig, cand) in neg or (cand, orig) in neg or orig == cand
                                    or is_pos(orig, cand) or is_pos(cand, orig)):
            orig = sampled_gids[gid_ind % n]
            gid_ind += 1
            cand = rng.choice(negative_gids, size=1)[0]
        neg.add((orig, cand))
    ordered = {k: 0 for k in neg}
    return np.asarray(list(ordered.keys()), np.int64).reshape(-1, 2)


class PairRecord(object):
    """utils/data/graph.py:16-50 GraphPair, minus the graph objects."""
    __slots__ = ('true_label', 'link_pred', 'gids')

    def __init__(self, true_label, gids):
        self.true_label, self.gids, self.link_pred = true_label, gids, None


class BatchData(object):
    def __init__(self, batch_gids, dataset, sampled_gids=None, curr_sample_edge_type=None, is_train=True,
                 ignore_pairs=False, enforce_negative_sampling=True, unique_graphs=True, subgraph=None,
                 merge_graphs=True):
        flags = get_flags()
        self.dataset = dataset
        self.is_train = is_train
        if isinstance(batch_gids, torch.Tensor):
            batch_gids = batch_gids.cpu().detach().numpy().astype('int')
        self.batch_gids = np.asarray(batch_gids)
        self.interaction_combo_nxgraph = dataset.interaction_combo_nxgraph
        self.unique_graphs = unique_graphs
        self.positive_pair_gids = self.batch_gids
        self.ignore_pairs = ignore_pairs
        if flags.negative_sample and is_train:
            assert sampled_gids is not None
            self.sampled_gids = sampled_gids
            neg = sample_negative_pairs(dataset, self.positive_pair_gids, sampled_gids,
                                        flags.num_negative_samples, enforce_negative=enforce_negative_sampling,
                                        amongst_same_graphs=getattr(flags, 'enforce_sampling_amongst_same_graphs',
                                                                    True))
            if len(neg) > 0:
                self.negative_pair_gids = neg
                self.batch_gids = np.concatenate((self.batch_gids, neg))
        # ---- labels (src/batch.py:117-130)
        self.pair_list = []
        if not ignore_pairs:
            for g1, g2 in self.batch_gids.tolist():
                l = dataset.look_up_label(g1, g2)
                self.pair_list.append(PairRecord(0 if l is None else l, (g1, g2)))
        # ---- merged molecule graph of the batch's unique drugs, built on the device
        order = unique_graphs_in_order(self.batch_gids)
        self.merge_data = {'gids_to_batch_ind': order}
        if merge_graphs:
            rows = np.asarray([dataset.gs_map[g] for g in order.keys()], np.int64)
            m = MergedGraph(dataset.packed, rows)
            self.merge_data.update({'merge': m, 'ind_list': m.ind_list(),
                                    'graph_sizes': np.diff(m.seg_ptr_host),
                                    'dataset_rows': torch.as_tensor(rows).to(dataset.device)})
        self.batch_interaction_inds = [dataset.gs_map[g] for g in self.batch_gids.flatten().tolist()]
        self.merge_higher_level = {}
        self._y = None
        self._pair_rows = None

    # -- device views used by the layers ---------------------------------------------
    def y_true_device(self, as_int=False):
        if self._y is None:
            y = np.asarray([p.true_label for p in self.pair_list])
            self._y = (torch.as_tensor(y.astype(np.float32)).to(self.dataset.device, non_blocking=True),
                       torch.as_tensor(y.astype(np.int32)).to(self.dataset.device, non_blocking=True))
        return self._y[1] if as_int else self._y[0]

    def pair_rows_device(self, n_rows, higher=True, unique=True):
        """[P,2] int32 rows of the scored embeddings (layers_link_pred.py:46-54) and the
        entry CSR used by the decoder's backward."""
        if self._pair_rows is None:
            if not unique:
                ids = np.arange(self.batch_gids.size).reshape(self.batch_gids.shape)
            elif higher:
                ids = np.asarray(self.batch_interaction_inds, np.int64).reshape(-1, 2)
            else:
                m = self.merge_data['gids_to_batch_ind']
                ids = np.asarray([m[g] for g in self.batch_gids.flatten().tolist()], np.int64).reshape(-1, 2)
            dev = self.dataset.device
            ids_dev = torch.as_tensor(ids.astype(np.int32)).to(dev, non_blocking=True)
            self._pair_rows = (ids_dev, entry_csr(ids, n_rows, dev))
        return self._pair_rows

    def assign_link_preds(self, pair_preds):
        """The reference calls .item() per pair (128 device syncs, utils/data/graph.py:35-39);
        predictions stay on the device until `link_preds()` is asked for."""
        self._preds = pair_preds.detach()

    def link_preds(self):
        p = self._preds.cpu().numpy()
        for rec, v in zip(self.pair_list, p):
            rec.link_pred = float(v) if v.ndim == 0 else (float(v[0]) if v.shape[0] == 1 else v)
        return p

    def restore_interaction_nxgraph(self):
        """src/batch.py:152-162: cut the autograd history of init_x after the step."""
        ig = self.dataset.interaction_combo_nxgraph
        if isinstance(ig.init_x, torch.Tensor):
            ig.init_x = ig.init_x.detach()
