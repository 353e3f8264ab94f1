"""Host-side dataset view the path consumes: what utils/data/dataset.py:47 BiGNNDataset and
:27 TorchBiGNNDataset expose to the hot path (gs_map, pairs, train pair tensor, the train
interaction graph), backed by flat arrays resident in HBM instead of networkx objects."""
import numpy as np
import torch

from .graph import PackedGraphs, InteractionGraph


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
        if pair_keys is not None:
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

    def __len__(self):
        return self.train_pairs.shape[0]

    def __getitem__(self, idx):
        return self.data_items[idx]

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
            self._edge_set = set(zip(g.row_host.tolist(), g.col_host.tolist()))
        return self._edge_set
