"""Host-side dataset view the path consumes: what utils/data/dataset.py:47 BiGNNDataset and
:27 TorchBiGNNDataset expose to the hot path (gs_map, pairs, train pair tensor, the train
interaction graph), backed by flat arrays resident in HBM instead of networkx objects."""
import numpy as np
import torch

from .graph import PackedGraphs, InteractionGraph

# above this many pairs / edges the Python dict / set of the reference (utils/data/dataset.py:394-403,
# src/batch.py:73) is replaced by binary search over sorted int64 keys: same answers, O(log E) per query,
# no per-call O(E) host work ("fine at 28 k, catastrophic at 20 M")
BIG_TABLE = 2_000_000


class _SortedKeySet(object):
    """`(a, b) in s` over sorted keys a*n + b."""

    def __init__(self, keys_sorted, n):
        self.keys, self.n = keys_sorted, int(n)

    def __contains__(self, ab):
        k = int(ab[0]) * self.n + int(ab[1])
        i = int(np.searchsorted(self.keys, k))
        return i < self.keys.shape[0] and int(self.keys[i]) == k

    def __len__(self):
        return int(self.keys.shape[0])


class _PairTable(object):
    """gid pair -> label for datasets too large for a dict of tuples."""

    def __init__(self, gs_map, n, pair_keys, pair_labels):
        gids = np.asarray(pair_keys, np.int64)
        lut_keys = np.fromiter(gs_map.keys(), np.int64, len(gs_map))
        lut_vals = np.fromiter(gs_map.values(), np.int64, len(gs_map))
        order = np.argsort(lut_keys)
        lut_keys, lut_vals = lut_keys[order], lut_vals[order]
        idx = np.minimum(np.searchsorted(lut_keys, gids), max(len(lut_keys) - 1, 0))
        if gids.size and not np.array_equal(lut_keys[idx], gids):
            raise KeyError('pair table: {} pair entries name a gid that is not in the dataset'.format(
                int((lut_keys[idx] != gids).sum())))
        rows = lut_vals[idx]
        key = rows[:, 0] * n + rows[:, 1]
        order = np.argsort(key, kind='stable')
        self.keys = key[order]
        self.labels = np.asarray(pair_labels)[order]
        self.gs_map, self.n = gs_map, int(n)

    def get(self, ab, default=None):
        a, b = self.gs_map.get(int(ab[0])), self.gs_map.get(int(ab[1]))
        if a is None or b is None:
            return default
        k = a * self.n + b
        i = int(np.searchsorted(self.keys, k))
        if i < self.keys.shape[0] and int(self.keys[i]) == k:
            return int(self.labels[i])
        return default

    def __len__(self):
        return int(self.keys.shape[0])


class BiGNNData(object):
    def __init__(self, gids, atom_ptr, nbr_ptr, nbr_idx, x, ddi_row, ddi_col, train_pairs,
                 pair_keys=None, pair_labels=None, num_labels=2, device='cuda:0', edge_types=None):
        self.device = torch.device(device)
        self.packed = PackedGraphs(gids, atom_ptr, nbr_ptr, nbr_idx, x, self.device)
        self.gs_map = self.packed.gs_map
        self.id_map = {i: g for g, i in self.gs_map.items()}
        self.N = self.packed.N
        self.num_node_feat = self.packed.num_node_feat
        self.num_labels = int(num_labels)
        self.interaction_num_node_feat = None
        self.num_hyper_edge_feat = 0
        self.interaction_combo_nxgraph = InteractionGraph(self.N, ddi_row, ddi_col, self.device)
        # per-edge-type interaction graphs (utils/data/dataset.py:105-115), DrugCombo: synergy / antagonism
        self.interaction_nxgraphs = {}
        if edge_types:
            for name in sorted(edge_types):
                r, c = edge_types[name]
                self.interaction_nxgraphs[name] = InteractionGraph(self.N, r, c, self.device)
            self.num_hyper_edge_feat = len(edge_types) + 1        # + the 'none' type of the added self loops
            # all edge types as ONE block-diagonal graph (type t occupies rows/cols [t*N, (t+1)*N)): the
            # per-type message passing of NodeModelAggrByEdge then runs as a single batched launch
            rows = np.concatenate([np.asarray(edge_types[n][0], np.int64) + t * self.N
                                   for t, n in enumerate(sorted(edge_types))])
            cols = np.concatenate([np.asarray(edge_types[n][1], np.int64) + t * self.N
                                   for t, n in enumerate(sorted(edge_types))])
            self.interaction_stack = InteractionGraph(len(edge_types) * self.N, rows, cols, self.device)
        # sorted (N*row+col) keys of the train graph for O(log E) membership tests
        self._edge_keys = np.sort(np.asarray(self.interaction_combo_nxgraph.row_host, np.int64) * self.N +
                                  np.asarray(self.interaction_combo_nxgraph.col_host, np.int64))
        self._edge_set = None
        self.train_pairs = np.asarray(train_pairs, np.int64)
        self.data_items = torch.as_tensor(self.train_pairs)        # sorted pair tensor (dataset.py:34)
        self.pairs = {}
        if pair_keys is not None and len(pair_keys) > BIG_TABLE:
            self.pairs = _PairTable(self.gs_map, self.N, pair_keys, pair_labels)
        elif pair_keys is not None:
            for (a, b), l in zip(np.asarray(pair_keys).tolist(), np.asarray(pair_labels).tolist()):
                self.pairs[(a, b)] = int(l)
        else:
            for a, b in self.train_pairs.tolist():
                self.pairs[(a, b)] = 1
        self.dataset = self            # reference code reaches both through `.dataset`

    @classmethod
    def from_npz(cls, path_or_npz, device='cuda:0'):
        z = np.load(path_or_npz) if isinstance(path_or_npz, str) else path_or_npz
        files = z.files if hasattr(z, 'files') else list(z.keys())
        x = z['x_u8'] if 'x_u8' in files else z['x']
        et = None
        names = [k[len('etype_row/'):] for k in files if k.startswith('etype_row/')]
        if names:
            et = {n: (z['etype_row/' + n], z['etype_col/' + n]) for n in names}
        return cls(z['gids'], z['atom_ptr'], z['nbr_ptr'], z['nbr_idx'], np.asarray(x, np.float32),
                   z['ddi_row'], z['ddi_col'], z['train_pairs'],
                   z['pair_keys'] if 'pair_keys' in files else None,
                   z['pair_labels'] if 'pair_labels' in files else None,
                   int(z['num_labels']) if 'num_labels' in files else 2, device, et)

    def init_interaction_graph_feats(self, init_method, d_init=64, feats=None, feat_size=None):
        """Fixed drug features for models WITHOUT a lower level (utils/data/dataset.py:176-198; src/load_data.py:
        150-153): 'rand_init' = xavier-normal [N, d_init] (gain of relu) drawn from torch's global RNG as the reference
        does, 'ones_init', or given `feats` ('graph_feats': the ECFP fingerprints the reference keeps in a klepto
        blob).  They become the interaction graph's node features (`init_x`), as `init_interaction_graph_embds` does
        (utils/data/dataset.py:166-175)."""
        if feats is not None:
            x = torch.as_tensor(np.asarray(feats, np.float32))
        elif 'rand_init' in init_method:
            x = torch.nn.init.xavier_normal_(torch.empty(self.N, d_init), gain=torch.nn.init.calculate_gain('relu'))
        elif 'ones_init' in init_method:
            x = torch.ones(self.N, d_init)
        else:
            raise NotImplementedError('init_embds={!r} needs the drug features passed in as `feats`'.format(init_method))
        self.graph_feats = x.to(self.device)
        self.interaction_num_node_feat = int(x.shape[1])
        self.interaction_combo_nxgraph.init_x = self.graph_feats
        return self.graph_feats

    def __len__(self):
        return self.train_pairs.shape[0]

    def __getitem__(self, idx):
        return self.data_items[idx]

    # ---- vectorised host lookups (large pair batches: engine_lower.LowerOnlyEngine, scaled benchmarks)
    def rows_of_gids(self, gids):
        """gs_map over an int64 array of gids."""
        if getattr(self, '_gid_lut', None) is None:
            keys = np.fromiter(self.gs_map.keys(), np.int64, len(self.gs_map))
            vals = np.fromiter(self.gs_map.values(), np.int64, len(self.gs_map))
            o = np.argsort(keys)
            self._gid_lut = (keys[o], vals[o])
        keys, vals = self._gid_lut
        g = np.asarray(gids, np.int64)
        i = np.minimum(np.searchsorted(keys, g), keys.shape[0] - 1)
        if g.size and not np.array_equal(keys[i], g):
            raise KeyError('unknown gid in a pair batch')
        return vals[i]

    def edge_keys_sorted(self):
        """sorted N*row+col keys of the train interaction graph (both orientations)."""
        return self._edge_keys

    def labels_of_pairs(self, gid_pairs):
        """look_up_label over an array of pairs (either orientation; 0 when unknown)."""
        p = np.asarray(gid_pairs, np.int64)
        if p.shape[0] == 0:
            return np.zeros(0, np.int64)
        if isinstance(self.pairs, _PairTable):
            keys, labels = self.pairs.keys, self.pairs.labels
        else:
            # the dict of tuples as a sorted key table, built once (a Python loop with two dict probes per pair holds
            # the GIL for 85 ms per 65 536-pair batch: the host threads of engine_lower.PairPrefetcher then serialise)
            if getattr(self, '_pairs_tab', None) is None or self._pairs_tab[2] != len(self.pairs):
                pk = np.asarray(list(self.pairs.keys()), np.int64).reshape(-1, 2)
                pl = np.asarray(list(self.pairs.values()), np.int64)
                rows = self.rows_of_gids(pk.reshape(-1)).reshape(-1, 2)
                key = rows[:, 0] * self.N + rows[:, 1]
                o = np.argsort(key, kind='stable')
                self._pairs_tab = (key[o], pl[o], len(self.pairs))
            keys, labels = self._pairs_tab[0], self._pairs_tab[1]
        if keys.shape[0] == 0:
            return np.zeros(p.shape[0], np.int64)
        rows = self.rows_of_gids(p.reshape(-1)).reshape(-1, 2)
        out = np.zeros(p.shape[0], np.int64)
        for a, b in ((0, 1), (1, 0)):
            k = rows[:, a] * self.N + rows[:, b]
            i = np.minimum(np.searchsorted(keys, k), keys.shape[0] - 1)
            hit = keys[i] == k
            out = np.where((out == 0) & hit, labels[i], out)
        return out

    def look_up_label(self, gid1, gid2):
        """utils/data/dataset.py:394-403 (None instead of ValueError when unknown)."""
        l = self.pairs.get((gid1, gid2))
        if l is None:
            l = self.pairs.get((gid2, gid1))
        return l

    def edge_set(self):
        """set(nx.edges) of the train interaction graph as (row,row) tuples, both
        orientations (src/batch.py:73 builds it per call; it never changes)."""
        if self._edge_set is None:
            g = self.interaction_combo_nxgraph
            if self._edge_keys.shape[0] > BIG_TABLE:
                self._edge_set = _SortedKeySet(self._edge_keys, self.N)
            else:
                self._edge_set = set(zip(g.row_host.tolist(), g.col_host.tolist()))
        return self._edge_set
