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
