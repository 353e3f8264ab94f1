"""Packed dataset format of the path: the flat arrays `BiGNNData` keeps in HBM, built
once from per-graph edge lists and feature matrices -- what the reference holds as networkx objects
plus float64 one-hot numpy (utils/data/load_raw_data.py:171-223, representation_node_feat.py:42-97)
and re-converts graph by graph every step (model/layers_util.py:100-166).  numpy only.

Layout (`.npz` keys): gids[N] i64 (ascending = gs_map order), atom_ptr[N+1] i32, nbr_ptr[sumA+1] i32,
nbr_idx[nnz] i32 (per-graph LOCAL neighbour ids; atoms and neighbours ascending = the sorted, coalesced,
symmetrised COO of create_edge_index), x_u8 or x [sumA, F], ddi_row/ddi_col (sorted symmetric train
interaction COO over drug rows), train_pairs[M,2] gids, pair_keys/pair_labels, num_labels and,
per interaction edge type, etype_row/<name>, etype_col/<name>.
"""
import numpy as np


def canonical_molecule_csr(edges, n):
    """sorted(edges) -> to_undirected -> coalesce (model/layers_util.py:148-166, PyG/torch-sparse
    semantics): per-atom ascending neighbour lists of the symmetrised, de-duplicated bond list."""
    e = np.asarray(edges, np.int64).reshape(-1, 2)
    if e.size == 0:
        return np.zeros(n + 1, np.int64), np.zeros(0, np.int64)
    key = np.unique(np.concatenate([e[:, 0] * n + e[:, 1], e[:, 1] * n + e[:, 0]]))
    row, col = key // n, key % n
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, row + 1, 1)
    return np.cumsum(ptr), col


def pack_dataset(gids, edge_lists, features, train_pairs, pair_labels=None, edge_types=None):
    """gids: graph ids; edge_lists[i]: [m_i,2] bonds of graph i over local atom ids; features[i]:
    [n_i,F] node features; train_pairs: [M,2] gid pairs of the train interaction graph;
    edge_types: optional {name: [m,2] gid pairs}.  Returns the dict `BiGNNData.from_npz` takes."""
    order = np.argsort(np.asarray(gids, np.int64), kind='stable')        # gs_map order = ascending gid
    gids = np.asarray(gids, np.int64)[order]
    row_of = {int(g): i for i, g in enumerate(gids)}
    atom_ptr, nbr_ptr, cols, xs = [0], [0], [], []
    for i in order:
        x = np.asarray(features[i])
        n = x.shape[0]
        ptr, col = canonical_molecule_csr(edge_lists[i], n)
        nbr_ptr.extend((nbr_ptr[-1] + ptr[1:]).tolist())
        cols.append(col)
        xs.append(x)
        atom_ptr.append(atom_ptr[-1] + n)
    x = np.concatenate(xs, 0)
    N = len(gids)

    def sym_coo(pairs):
        p = np.asarray(pairs, np.int64).reshape(-1, 2)
        a = np.asarray([row_of[int(g)] for g in p[:, 0]], np.int64)
        b = np.asarray([row_of[int(g)] for g in p[:, 1]], np.int64)
        key = np.unique(np.concatenate([a * N + b, b * N + a]))
        return (key // N).astype(np.int32), (key % N).astype(np.int32)

    tp = np.asarray(sorted(map(tuple, np.asarray(train_pairs, np.int64).tolist())), np.int64)
    r, c = sym_coo(tp)
    out = dict(gids=gids, atom_ptr=np.asarray(atom_ptr, np.int32), nbr_ptr=np.asarray(nbr_ptr, np.int32),
               nbr_idx=np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32),
               ddi_row=r, ddi_col=c, train_pairs=tp, pair_keys=tp,
               pair_labels=np.asarray(pair_labels if pair_labels is not None else np.ones(len(tp)), np.int8),
               num_labels=np.int64(len(set(np.asarray(pair_labels).tolist())) if pair_labels is not None else 2))
    out['x_u8' if np.all((x == 0) | (x == 1)) else 'x'] = x.astype(np.uint8 if np.all((x == 0) | (x == 1)) else np.float32)
    if edge_types:
        for name, pairs in edge_types.items():
            er, ec = sym_coo(pairs)
            out['etype_row/' + name], out['etype_col/' + name] = er, ec
    return out


def save_npz(path, packed):
    np.savez_compressed(path, **packed)
