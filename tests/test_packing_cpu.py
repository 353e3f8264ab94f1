"""Packed dataset format (no GPU)."""
import os

import numpy as np

import bignn_b200  # noqa: F401


def test_packing_round_trip_equals_golden_pack(golden_dir, drugbank):
    """the product-side packer reproduces the reference-generated packed arrays from edge lists."""
    from bignn_b200.packing import pack_dataset
    z = np.load(os.path.join(golden_dir, 'drugbank_packed.npz'))
    edges, feats = [], []
    for i in range(drugbank.N):
        und, n = drugbank.mol_undirected_edges(i)
        edges.append(und[:, ::-1])                         # either orientation must do
        feats.append(drugbank.x[drugbank.atom_ptr[i]:drugbank.atom_ptr[i + 1]])
    packed = pack_dataset(z['gids'], edges, feats, z['train_pairs'])
    for k in ('gids', 'atom_ptr', 'nbr_ptr', 'nbr_idx', 'x_u8', 'ddi_row', 'ddi_col', 'train_pairs'):
        assert np.array_equal(packed[k], z[k]), k
