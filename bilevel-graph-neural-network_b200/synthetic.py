"""Synthetic workloads of the shapes BASELINE.json names: molecule-like graphs
(S-mol) and a degree-skewed drug-drug interaction graph (S-ddi), as the flat arrays BiGNNData /
the oracle's PackedDataset take.  Deterministic in (shape, seed); numpy only."""
import numpy as np

FEATURE_GROUPS = (27, 2, 2, 2, 7, 9)      # six one-hot groups, 49 columns (DrugBank's F_in)


def molecule_graphs(G, mean_atoms=28.0, seed=0, min_atoms=2, max_atoms=457, sigma=0.45, chord_frac=0.08,
                    groups=FEATURE_GROUPS, backbone=False):
    """G molecule-like graphs: sizes ~ clipped log-normal rescaled to `mean_atoms`; edges = random
    spanning tree + round(chord_frac*n) chords (deduplicated, no self loops); features = one-hot
    groups with a Zipf-ish category skew.  backbone=True (protein-like representation graphs, SURVEY 8d S-mol for C5):
    the tree is the residue chain i -- i+1 and the chords are random contacts (chord_frac = 3 gives mean degree ~8).
    Returns atom_ptr, nbr_ptr, nbr_idx (local ids, sorted), x."""
    rng = np.random.default_rng(seed)
    n = np.exp(rng.normal(np.log(26.0), sigma, G))
    n = np.clip(np.round(n * (mean_atoms / n.mean())), min_atoms, max_atoms).astype(np.int64)
    atom_ptr = np.concatenate([[0], np.cumsum(n)])
    A = int(atom_ptr[-1])
    # spanning tree: atom i>0 of a graph attaches to a random earlier atom of the same graph (or to i-1: a chain)
    local = np.arange(A) - np.repeat(atom_ptr[:-1], n)
    gid = np.repeat(np.arange(G), n)
    child = np.nonzero(local > 0)[0]
    parent_local = local[child] - 1 if backbone else (rng.random(child.shape[0]) * local[child]).astype(np.int64)
    src = [local[child]]
    dst = [parent_local]
    eg = [gid[child]]
    # chords
    k = np.round(chord_frac * n).astype(np.int64)
    cg = np.repeat(np.arange(G), k)
    a = (rng.random(cg.shape[0]) * n[cg]).astype(np.int64)
    b = (rng.random(cg.shape[0]) * n[cg]).astype(np.int64)
    keep = a != b
    src.append(a[keep]); dst.append(b[keep]); eg.append(cg[keep])
    s = np.concatenate(src); d = np.concatenate(dst); g = np.concatenate(eg)
    # symmetrise, globalise, sort, dedupe
    gs = np.concatenate([atom_ptr[g] + s, atom_ptr[g] + d])
    gd = np.concatenate([atom_ptr[g] + d, atom_ptr[g] + s])
    key = np.unique(gs * A + gd)
    row, col = key // A, key % A
    nbr_ptr = np.zeros(A + 1, np.int64)
    np.add.at(nbr_ptr, row + 1, 1)
    nbr_ptr = np.cumsum(nbr_ptr)
    nbr_idx = col - atom_ptr[np.searchsorted(atom_ptr, col, side='right') - 1]
    F = int(sum(groups))
    x = np.zeros((A, F), np.uint8)
    off = 0
    for w in groups:
        p = 1.0 / np.arange(1, w + 1) ** 1.5
        c = rng.choice(w, size=A, p=p / p.sum())
        x[np.arange(A), off + c] = 1
        off += w
    return atom_ptr.astype(np.int32), nbr_ptr.astype(np.int32), nbr_idx.astype(np.int32), x


def interaction_graph(N, M, seed=0, skew=0.8):
    """M undirected drug pairs over N drugs, p(v) ~ rank^-skew, no self loops, deduplicated.
    Returns the undirected pair list (a<b, sorted) and the symmetric sorted directed COO."""
    rng = np.random.default_rng(seed + 1000003)
    p = 1.0 / np.arange(1, N + 1) ** skew
    cdf = np.cumsum(p / p.sum())
    perm = rng.permutation(N)
    pairs = np.zeros((0, 2), np.int64)
    need = M
    while need > 0:
        k = int(need * 1.3) + 16
        a = perm[np.minimum(np.searchsorted(cdf, rng.random(k)), N - 1)]
        b = perm[np.minimum(np.searchsorted(cdf, rng.random(k)), N - 1)]
        keep = a != b
        lo, hi = np.minimum(a[keep], b[keep]), np.maximum(a[keep], b[keep])
        key = np.unique(np.concatenate([pairs[:, 0] * N + pairs[:, 1], lo * N + hi]))
        if key.shape[0] > M:
            key = np.sort(rng.choice(key, M, replace=False))
        pairs = np.stack([key // N, key % N], 1)
        need = M - pairs.shape[0]
    key = np.unique(np.concatenate([pairs[:, 0] * N + pairs[:, 1], pairs[:, 1] * N + pairs[:, 0]]))
    return pairs, key // N, key % N


def bignn_workload(N, M, mean_atoms=28.0, seed=0, groups=FEATURE_GROUPS, max_atoms=457, edge_type_fracs=None,
                   min_atoms=2, chord_frac=0.08, backbone=False):
    """A full Bi-GNN dataset dict (keys as tests/golden/drugbank_packed.npz).  With
    `edge_type_fracs` (e.g. {'synergy': 0.68, 'antagonism': 0.32}) every interaction gets one
    edge type: pair labels become 1..T (0 = sampled negative) and per-type directed COOs are added
    as `etype_row/<name>`, `etype_col/<name>` (the DrugCombo layout, utils/data/dataset.py:82-115)."""
    atom_ptr, nbr_ptr, nbr_idx, x = molecule_graphs(N, mean_atoms, seed, groups=groups, max_atoms=max_atoms,
                                                    min_atoms=min_atoms, chord_frac=chord_frac, backbone=backbone)
    pairs, row, col = interaction_graph(N, M, seed)
    gids = np.arange(N, dtype=np.int64) + 1000          # gids are labels, not row numbers
    train_pairs = gids[pairs]
    out = dict(gids=gids, atom_ptr=atom_ptr, nbr_ptr=nbr_ptr, nbr_idx=nbr_idx, x_u8=x,
               ddi_row=row.astype(np.int32), ddi_col=col.astype(np.int32), train_pairs=train_pairs,
               pair_keys=train_pairs, pair_labels=np.ones(train_pairs.shape[0], np.int8),
               num_labels=np.int64(2))
    if edge_type_fracs:
        rng = np.random.default_rng(seed + 77)
        names = sorted(edge_type_fracs)
        p = np.asarray([edge_type_fracs[n] for n in names], np.float64)
        t = rng.choice(len(names), size=pairs.shape[0], p=p / p.sum())
        out['pair_labels'] = (t + 1).astype(np.int8)
        out['num_labels'] = np.int64(len(names))
        for i, n in enumerate(names):
            pr = pairs[t == i]
            key = np.unique(np.concatenate([pr[:, 0] * N + pr[:, 1], pr[:, 1] * N + pr[:, 0]]))
            out['etype_row/' + n] = (key // N).astype(np.int32)
            out['etype_col/' + n] = (key % N).astype(np.int32)
    return out


WORKLOADS = {
    # name: (N drugs, M undirected DDI edges, mean atoms, feature groups)
    'drugbank_shape': dict(N=1309, M=28751, mean_atoms=28.1, groups=FEATURE_GROUPS),
    # DrugCombo (App. D): 3 242 drugs in the interaction graphs, 29.3 atoms, synergy 34 355 +
    # antagonism 15 908 undirected edges, one-hot width <= 40
    'drugcombo_shape': dict(N=3242, M=50263, mean_atoms=29.3, groups=(22, 2, 2, 2, 6, 6),
                            edge_type_fracs={'synergy': 34355.0, 'antagonism': 15908.0}),
    # BASELINE config 4 / SURVEY C4: 200 k drugs, 20 M undirected DDI edges (40 M nnz), 64-dim upper level
    'ddi_scaled': dict(N=200_000, M=20_000_000, mean_atoms=30.0, groups=FEATURE_GROUPS),
    # the same shape at 1/10 (CPU baseline of the scaled configuration; quick multi-GPU checks)
    'ddi_scaled_small': dict(N=20_000, M=2_000_000, mean_atoms=30.0, groups=FEATURE_GROUPS),
    # BASELINE config 3 / SURVEY C3: 1 M molecule graphs x ~30 atoms (~65 directed bonds each); lower-level-only model
    # (pairs drawn from a sparse 2 M-edge interaction list; the model never reads the interaction graph)
    'mol_1m': dict(N=1_000_000, M=2_000_000, mean_atoms=30.0, groups=FEATURE_GROUPS),
    'mol_1m_small': dict(N=100_000, M=200_000, mean_atoms=30.0, groups=FEATURE_GROUPS),
    # BASELINE config 5 / SURVEY C5: 50 k protein-like representation graphs x ~400 residues (chain + random contacts,
    # mean degree ~8; clipped to [50, 1000] residues) + a 5 M-edge interaction graph
    'ppi_50k': dict(N=50_000, M=5_000_000, mean_atoms=400.0, groups=FEATURE_GROUPS, min_atoms=50, max_atoms=1000,
                    chord_frac=3.0, backbone=True),
    'ppi_50k_small': dict(N=5_000, M=500_000, mean_atoms=400.0, groups=FEATURE_GROUPS, min_atoms=50, max_atoms=1000,
                          chord_frac=3.0, backbone=True),
}


def cached_workload(name, seed=0, cache_dir=None, writer=True, wait_s=600.0):
    """bignn_workload(**WORKLOADS[name]) through an .npz cache in the temp directory (the 20 M-edge graph
    takes about a minute to draw; every rank of a multi-GPU job and the CPU baseline need the same arrays).
    writer=False (ranks other than local rank 0): wait for the writer's file instead of drawing the same graph
    N times at once; falls back to drawing it after `wait_s`."""
    import os
    import tempfile
    import time
    w = WORKLOADS[name]
    if w['N'] < 10_000:
        return bignn_workload(seed=seed, **w)
    path = os.path.join(cache_dir or tempfile.gettempdir(), 'bignn_synth_{}_{}.npz'.format(name, seed))

    def load():
        z = np.load(path)
        return {k: z[k] for k in z.files}
    t0 = time.time()
    while True:
        if os.path.exists(path):
            try:
                return load()
            except Exception:
                pass
        if writer or time.time() - t0 > wait_s:
            break
        time.sleep(0.5)
    out = bignn_workload(seed=seed, **w)
    tmp = '{}.{}.tmp.npz'.format(path, os.getpid())
    np.savez(tmp, **out)
    os.replace(tmp, path)
    return out
