"""Known-answer micro cases for the CPU oracle and the packer: hand-made graphs whose
results can be written down -- path P3, star, two components, isolated node, duplicate / self-loop
edge input."""
import numpy as np
import torch

import bignn_b200  # noqa: F401
from bignn_b200.packing import canonical_molecule_csr
from oracle import bignn_oracle as O


def test_coalesce_drops_duplicates_keeps_order():
    # duplicates, both orientations and unsorted input collapse to the sorted symmetric COO
    ei = O.coalesce_undirected([(2, 1), (0, 1), (1, 0), (1, 2), (0, 1)], 3)
    assert ei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    ptr, col = canonical_molecule_csr([(2, 1), (0, 1), (1, 0), (1, 2), (0, 1)], 3)
    assert ptr.tolist() == [0, 1, 3, 4] and col.tolist() == [1, 0, 2, 1]
    # a self loop survives canonicalisation (the convolutions remove it: App. A.1/A.2)
    ei = O.coalesce_undirected([(0, 0), (0, 1)], 2)
    assert ei.tolist() == [[0, 0, 1], [0, 1, 0]]
    assert O.coalesce_undirected([], 4).shape == (2, 0)


def test_gin_and_gcn_on_path_p3_by_hand():
    # P3: 0 - 1 - 2, features x = [1, 10, 100]
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    x = torch.tensor([[1.0], [10.0], [100.0]])
    P = {'l.conv.eps': torch.zeros(1), 'l.conv.nn.0.weight': torch.ones(1, 1), 'l.conv.nn.0.bias': torch.zeros(1),
         'l.conv.nn.2.weight': torch.ones(1, 1), 'l.conv.nn.2.bias': torch.zeros(1)}
    out = O.gin_conv(x, ei, P, 'l', 'identity')
    assert out.view(-1).tolist() == [11.0, 111.0, 110.0]            # x_i + sum of neighbours
    Pg = {'l.conv.weight': torch.ones(1, 1), 'l.conv.bias': torch.zeros(1)}
    out = O.gcn_conv(x, ei, Pg, 'l').view(-1)
    d = torch.tensor([2.0, 3.0, 2.0]).pow(-0.5)                      # degrees incl. the self loop
    want = torch.stack([d[0] * d[0] * 1 + d[0] * d[1] * 10,
                        d[1] * d[0] * 1 + d[1] * d[1] * 10 + d[1] * d[2] * 100,
                        d[2] * d[1] * 10 + d[2] * d[2] * 100])
    assert torch.allclose(out, want, rtol=1e-6)


def test_star_two_components_and_isolated_node():
    # star centred at 0 with leaves 1..3, plus component 4-5, plus isolated node 6
    e = [(0, 1), (0, 2), (0, 3), (4, 5)]
    ei = torch.from_numpy(O.coalesce_undirected(e, 7))
    x = torch.arange(1.0, 8.0).view(7, 1)
    agg = O._scatter_rows(x.index_select(0, ei[0]), ei[1], 7).view(-1)
    assert agg.tolist() == [9.0, 1.0, 1.0, 1.0, 6.0, 5.0, 0.0]
    Pg = {'l.conv.weight': torch.ones(1, 1), 'l.conv.bias': torch.zeros(1)}
    out = O.gcn_conv(x, ei, Pg, 'l').view(-1)
    assert abs(float(out[6]) - 7.0) < 1e-6                            # isolated: only its self loop, deg 1
    # mean readout over graphs {0..3}, {4,5}, {6}
    batch = torch.tensor([0, 0, 0, 0, 1, 1, 2])
    assert O.readout([x], batch, 3, 'avg_pool').view(-1).tolist() == [2.5, 5.5, 7.0]
    assert O.readout([x], batch, 3, 'sum').view(-1).tolist() == [10.0, 11.0, 7.0]


def test_chunk_schedule_small_cases():
    # src/train.py:52-71: pairs (0,1),(2,3),... plus the final (N-2,N-1); chunk = batch_size pairs
    ch = O.all_drug_chunks(list(range(7)), 2)
    assert [c.tolist() for c in ch] == [[[0, 1], [2, 3]], [[4, 5], [5, 6]]]
    assert O.unique_graphs_in_order(ch[1]) == [4, 5, 6]              # 5 appears twice, merged once
    ch = O.all_drug_chunks(list(range(6)), 64)         # 2*bs >= #pairs: chunk = #pairs // 2 = 1 pair
    assert [len(c) for c in ch] == [1, 1, 1]


def test_smallest_and_largest_molecule_of_drugbank(drugbank):
    """the smallest usable molecule (2 atoms, 1 bond) and the 457-atom maximum, through the oracle's
    merge and the host-side MergedGraph (CPU stand-in backend): indexing by hand for the former, consistency
    properties for the latter."""
    import bignn_b200 as B
    from tests import fake_backend
    ds = drugbank
    sizes = np.diff(ds.atom_ptr)
    small, big = int(np.argmin(sizes)), int(np.argmax(sizes))
    assert sizes[small] == 2 and sizes[big] == 457
    m = O.merge_graphs(ds, [int(ds.gids[small]), int(ds.gids[big])])
    # graph 0: atoms 0,1 joined by one bond -> directed entries (0,1),(1,0); graph 1 starts at node 2
    assert m['edge_index'][:, :2].tolist() == [[0, 1], [1, 0]]
    assert m['ind_list'].tolist() == [[0, 2], [2, 459]] and m['edge_ind_list'][0].tolist() == [0, 2]
    assert m['batch'].tolist() == [0, 0] + [1] * 457
    ei = m['edge_index'][:, 2:]
    assert ei.min() == 2 and ei.max() == 458                                   # offset by the first graph's atoms
    key = ei[0].astype(np.int64) * 459 + ei[1]
    assert np.all(np.diff(key) > 0)                                            # sorted, no duplicates
    assert set(map(tuple, ei.T.tolist())) == set(map(tuple, ei[::-1].T.tolist()))   # symmetric
    assert np.all(m['x'].sum(1) == 6) and set(np.unique(m['x'])) <= {0.0, 1.0}       # six one-hot groups per atom
    fake_backend.install()
    try:
        B.set_flags(B.make_flags(device='cpu'))
        import os
        from tests.conftest import GOLDEN
        data = B.BiGNNData.from_npz(os.path.join(GOLDEN, 'drugbank_packed.npz'), device='cpu')
        g = B.MergedGraph(data.packed, [small, big])
        assert np.array_equal(g.edge_index.numpy(), m['edge_index'].astype(np.int64))
        assert np.array_equal(g.batch.numpy(), m['batch'].astype(np.int64))
        assert np.array_equal(g.x.numpy(), m['x'])
        assert g.seg_ptr.tolist() == [0, 2, 459] and g.row_ptr[:3].tolist() == [0, 1, 2]
    finally:
        fake_backend.uninstall()
        B.set_flags(None)


def test_row_plan_items_and_hub_rows():
    """work-item plan of a skewed CSR (ops.RowPlan): ceil(deg/seg) items per row (>= 1), multi-item rows listed with
    the hub rows (> BIG_ITEMS items) LAST, as bignn_spmm_planned_rows_f32 expects."""
    from bignn_b200 import ops
    deg = np.asarray([0, 1, 32, 33, 64, 65, 5000, 31, 2049, 2048])
    ptr = np.concatenate([[0], np.cumsum(deg)])
    pl = ops.RowPlan(ptr, 'cpu', seg=32)
    items = np.maximum(1, -(-deg // 32))
    assert pl.n_items == int(items.sum()) and pl.item_ptr.tolist() == np.concatenate([[0], np.cumsum(items)]).tolist()
    assert pl.item_row.tolist() == np.repeat(np.arange(10), items).tolist()
    multi = pl.multi_rows.tolist()
    assert sorted(multi) == [3, 4, 5, 6, 8, 9] and pl.n_multi == 6
    assert pl.n_big == 2 and multi[-2:] == [6, 8]            # 157 and 65 items; row 9 has exactly 64 = not a hub
    assert multi[:4] == [3, 4, 5, 9]                          # the others in ascending order
