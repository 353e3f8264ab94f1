"""Device-resident graph containers of the path.

PackedGraphs      all molecule graphs of a dataset as flat arrays in HBM (uploaded once);
                  replaces the per-graph networkx objects + float64 one-hot numpy of
                  utils/data/dataset.py:47 / representation_node_feat.py:42-97 as the source
                  of the merged batches.
MergedGraph       a block-diagonal merged batch built ON THE DEVICE by bignn_merge_build;
                  replaces src/merged_graph.py:13 MergedGraphData (`x`, `edge_index`,
                  `batch` keep their reference meaning; int64 views are materialised lazily).
InteractionGraph  the upper-level drug-drug CSR (train edges), built once; replaces the
                  per-step nx -> Data conversion of model/layers_load_interaction_graph.py:11-21.
"""
import numpy as np
import torch

from . import _lib
from .ops import CSR


def _i32(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(device)


class PackedGraphs(object):
    def __init__(self, gids, atom_ptr, nbr_ptr, nbr_idx, x, device):
        self.device = torch.device(device)
        self.gids = np.asarray(gids, np.int64)
        self.gs_map = {int(g): i for i, g in enumerate(self.gids)}       # gid -> row (dataset.py:369)
        self.atom_ptr_host = np.ascontiguousarray(atom_ptr, dtype=np.int64)
        self.nbr_ptr_host = np.ascontiguousarray(nbr_ptr, dtype=np.int64)
        if self.atom_ptr_host[-1] >= 2 ** 31 or self.nbr_ptr_host[-1] >= 2 ** 31:
            raise ValueError('packed dataset exceeds int32 indexing; shard it by drug')
        self.N = len(self.gids)
        self.num_node_feat = int(x.shape[1])
        self.atom_ptr = _i32(atom_ptr, self.device)
        self.nbr_ptr = _i32(nbr_ptr, self.device)
        self.nbr_idx = _i32(nbr_idx, self.device)
        self.x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self.device)
        # directed-edge offset of every graph (host) for output sizing
        self.edge_ptr_host = self.nbr_ptr_host[self.atom_ptr_host]

    @classmethod
    def from_npz(cls, z, device):
        x = z['x_u8'] if 'x_u8' in z else z['x']
        return cls(z['gids'], z['atom_ptr'], z['nbr_ptr'], z['nbr_idx'], np.asarray(x, np.float32), device)

    def x_padded(self):
        """(x with the feature columns zero-padded to a multiple of 4, padded width): 16-byte aligned rows for the
        fused layer kernel's 128-bit gathers (one-time copy; 49 -> 52 columns on DrugBank)."""
        if getattr(self, '_x_pad', None) is None:
            F = self.num_node_feat
            Fp = (F + 3) // 4 * 4
            if Fp == F:
                self._x_pad = (self.x, F)
            else:
                xp = torch.zeros((self.x.shape[0], Fp), dtype=torch.float32, device=self.device)
                xp[:, :F] = self.x
                self._x_pad = (xp, Fp)
        return self._x_pad

    def sizes(self, rows):
        rows = np.asarray(rows, np.int64)
        n = self.atom_ptr_host[rows + 1] - self.atom_ptr_host[rows]
        e = self.edge_ptr_host[rows + 1] - self.edge_ptr_host[rows]
        return n, e


class MergedGraph(object):
    """Merged batch on the device.  `chunk_row_ptr` (optional) marks groups of graphs that
    form independent BatchNorm batches (the 128-graph chunks of src/train.py:62-71)."""

    def __init__(self, packed, rows, chunk_graph_ptr=None, with_x=True, pad_features=False):
        rows = np.asarray(rows, np.int64)
        self._src_x, self.feat_width = packed.x_padded() if pad_features else (packed.x, packed.num_node_feat)
        self.packed = packed
        self.G = int(rows.shape[0])
        n, e = packed.sizes(rows) if self.G else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        self.A, self.E = int(n.sum()), int(e.sum())
        self.seg_ptr_host = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        dev = packed.device
        self.rows = _i32(rows, dev)
        self.seg_ptr = torch.empty(self.G + 1, dtype=torch.int32, device=dev)
        self.edge_ptr = torch.empty(self.G + 1, dtype=torch.int32, device=dev)
        self.row_ptr = torch.empty(self.A + 1, dtype=torch.int32, device=dev)
        self.col_idx = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)[:self.E]
        self.batch_i32 = torch.empty(max(self.A, 1), dtype=torch.int32, device=dev)[:self.A]
        self.x = torch.empty((self.A, self.feat_width), dtype=torch.float32, device=dev) if with_x else None
        self.edge_attr = None                 # molecule graphs carry no edge features (layers.py:48-50)
        self._edge_index = None
        self._batch = None
        if chunk_graph_ptr is None:
            chunk_graph_ptr = np.asarray([0, self.G], np.int64)
        self.chunk_graph_ptr_host = np.asarray(chunk_graph_ptr, np.int64)
        self.S = len(self.chunk_graph_ptr_host) - 1
        self.chunk_row_ptr = _i32(self.seg_ptr_host[self.chunk_graph_ptr_host], dev)
        self.chunk_graph_ptr = _i32(self.chunk_graph_ptr_host, dev)
        self._ws_bytes = int(_lib.call('bignn_merge_build_workspace_bytes', self.G))
        self._ws = torch.empty(max(self._ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.build()
        self.csr = CSR(self.row_ptr, self.col_idx, self.A)

    def build(self, edge_index=None, batch64=None):
        """(Re)runs the device-side construction; capturable in a CUDA graph."""
        p = self.packed
        _lib.require_device(p.atom_ptr)
        _lib.call('bignn_merge_build', p.atom_ptr, p.nbr_ptr, p.nbr_idx, self._src_x, self.feat_width,
                  self.rows, self.G, self.seg_ptr, self.edge_ptr, self.row_ptr, self.col_idx, self.batch_i32,
                  self.x, edge_index, batch64, self.A, self.E, self._ws, self._ws_bytes)

    def fused_plan(self):
        """Static index helpers of the fused layer kernel: the chunk of every 128-row tile's first row, the chunk of
        every graph, and the number of (tile, chunk) statistics records."""
        if getattr(self, '_fused_plan', None) is None:
            dev = self.packed.device
            crp = self.seg_ptr_host[self.chunk_graph_ptr_host]
            n_tiles = max((self.A + 127) // 128, 1)
            tile0 = np.searchsorted(crp, np.arange(n_tiles, dtype=np.int64) * 128, side='right') - 1
            tile0 = np.clip(tile0, 0, max(self.S - 1, 0))
            gchunk = np.searchsorted(self.chunk_graph_ptr_host, np.arange(self.G, dtype=np.int64), side='right') - 1
            # first CSR entry of every tile (row_ptr[128 t], row_ptr[A] last): lets the kernel prefetch a tile's
            # neighbour ids without first waiting for its row pointers
            pos = torch.as_tensor(np.minimum(np.arange(n_tiles + 1, dtype=np.int64) * 128, self.A)).to(dev)
            tile_edge = self.row_ptr.index_select(0, pos).contiguous()
            self._fused_plan = dict(tile_chunk0=_i32(tile0, dev), graph_chunk=_i32(np.clip(gchunk, 0, max(self.S - 1, 0)), dev),
                                    tile_edge_ptr=tile_edge, n_tiles=int(n_tiles), records=int(n_tiles + self.S))
        return self._fused_plan

    @property
    def edge_index(self):
        """int64 [2,E] COO, lexicographically sorted -- the reference's `edge_index`."""
        if self._edge_index is None:
            self._edge_index = torch.empty((2, self.E), dtype=torch.int64, device=self.packed.device)
            self._batch = torch.empty(self.A, dtype=torch.int64, device=self.packed.device)
            self.build(self._edge_index, self._batch)
        return self._edge_index

    @property
    def batch(self):
        if self._batch is None:
            self.edge_index
        return self._batch

    @property
    def num_nodes(self):
        return self.A

    def ind_list(self):
        s = self.seg_ptr_host
        return [(int(s[i]), int(s[i + 1])) for i in range(self.G)]


class InteractionGraph(object):
    """Upper-level graph over drug rows: sorted directed COO -> int32 CSR in HBM."""

    def __init__(self, n_nodes, row, col, device, x=None):
        row = np.asarray(row, np.int64)
        col = np.asarray(col, np.int64)
        key = row * n_nodes + col
        if key.size and np.any(np.diff(key) <= 0):
            order = np.argsort(key, kind='stable')
            row, col = row[order], col[order]
            keep = np.concatenate([[True], np.diff(key[order]) > 0])
            row, col = row[keep], col[keep]
        self.n = int(n_nodes)
        self.row_host, self.col_host = row, col
        ptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=self.n))]).astype(np.int64)
        dev = torch.device(device)
        self.csr = CSR(_i32(ptr, dev), _i32(col, dev), self.n, row_ptr_host=ptr)
        self.bn_row_ptr = _i32([0, self.n], dev)     # the whole graph is one BatchNorm batch
        self.init_x = x                              # [n, D] node features (pooled drug embeddings)
        self.edge_attr = None
        self._edge_index = None

    @property
    def x(self):
        return self.init_x

    def number_of_nodes(self):
        return self.n

    @property
    def edge_index(self):
        if self._edge_index is None:
            self._edge_index = torch.as_tensor(np.stack([self.row_host, self.col_host])).to(self.csr.row_ptr.device)
        return self._edge_index


def entry_csr(ids_host, n_rows, device):
    """CSR over drug rows of the 2P (pair, side) entries that reference them, entries in
    ascending order -- the transpose of the decoder's row gather (stable counting sort on
    the host; P is the pair batch, 128 by default)."""
    flat = np.asarray(ids_host, np.int64).reshape(-1)
    order = np.argsort(flat, kind='stable')
    ptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.bincount(flat, minlength=n_rows), out=ptr[1:])
    return CSR(_i32(ptr, device), _i32(order, device), n_rows)


class RowPartition(object):
    """Contiguous row ranges of the interaction graph, one per rank, balanced by work (nnz + rows) --
    'edges partitioned by source drug' (SURVEY 8e).  Every rank's block is padded to `n_max` rows so that
    the per-layer exchange is one equal-count all-gather: drug row g of rank r lives at position
    r*n_max + (g - lo_r) of the gathered [world*n_max, D] matrix, and the local CSRs store their column
    ids in that position space (the map is monotone, so neighbour order -- and with it the summation
    order of every SpMM row -- is the one of the unpartitioned graph)."""

    def __init__(self, row_ptr_host, world):
        ptr = np.asarray(row_ptr_host, np.int64)
        n = ptr.shape[0] - 1
        work = ptr + np.arange(n + 1, dtype=np.int64)            # prefix of (nnz + 1) per row
        bounds = [0]
        for r in range(1, world):
            i = int(np.searchsorted(work, work[-1] * r / world))
            bounds.append(min(max(i, bounds[-1]), n))
        bounds.append(n)
        self.world = int(world)
        self.bounds = np.asarray(bounds, np.int64)
        self.n = n
        self.n_max = max(int(np.diff(self.bounds).max()), 1)
        self.n_pad = self.world * self.n_max

    def rows_of(self, rank):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    def pos(self, rows):
        """drug rows -> positions in the gathered matrix."""
        rows = np.asarray(rows, np.int64)
        r = np.searchsorted(self.bounds[1:], rows, side='right')
        r = np.minimum(r, self.world - 1)
        return r * self.n_max + rows - self.bounds[r]


class PartitionedInteractionGraph(object):
    """This rank's rows of the upper-level graph (see RowPartition): local int32 CSR with columns in the
    gathered position space, the cached deg^-1/2 of ALL nodes in that space, and the pooled drug
    embeddings `init_x` of the local rows.  Stands in for InteractionGraph on `batch.merge_higher_level`;
    the layers recognise it by `partitioned`."""
    partitioned = True

    def __init__(self, full, rank, world, group=None):
        ptr = np.asarray(full.csr._row_ptr_host, np.int64)
        self.part = RowPartition(ptr, world)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.full = full
        self.n = full.n                                          # rows of the whole batch (BatchNorm count)
        lo, hi = self.part.rows_of(rank)
        self.lo, self.hi, self.n_loc = lo, hi, hi - lo
        self.n_max, self.n_pad = self.part.n_max, self.part.n_pad
        self.row_offset = self.rank * self.n_max
        dev = full.csr.row_ptr.device
        lptr = ptr[lo:hi + 1] - ptr[lo]
        lcol = self.part.pos(full.col_host[ptr[lo]:ptr[hi]])
        self.csr = CSR(_i32(lptr, dev), _i32(lcol, dev), self.n_loc, row_ptr_host=lptr)
        self.csr.row_offset = self.row_offset
        # deg^-1/2 of every node, scattered to the position space (pad positions are never referenced)
        dinv = torch.zeros(self.n_pad, dtype=torch.float32, device=dev)
        dinv[torch.as_tensor(self.part.pos(np.arange(full.n))).to(dev)] = full.csr.dinv()
        self.csr._dinv = dinv
        self.init_x_full = None
        self.edge_attr = None
        self.bn_row_ptr = None                 # BatchNorm goes through bignn_bn_rows_* (batch = rows of all ranks)

    @property
    def init_x(self):
        return None if self.init_x_full is None else self.init_x_full[self.lo:self.hi]

    @init_x.setter
    def init_x(self, value):
        self.init_x_full = value

    @property
    def x(self):
        return self.init_x

    def number_of_nodes(self):
        return self.n
