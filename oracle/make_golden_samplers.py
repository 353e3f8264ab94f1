"""TEST INFRASTRUCTURE ONLY -- records the reference's NeighborSampler / EverythingSampler (src/sampler.py:51-107)
on DrugBank fold 1 after set_seed(8): six consecutive NeighborSampler(neighbor_size=5, batch_size=64) batches
(sampled drugs in order, induced pairs in order, sub-graph node order, the visit counter) the head of one
EverythingSampler batch, and four RandomSampler(sample_induced=True) batches -> tests/golden/bignn_samplers.npz.  Runs only where /root/reference exists.
Note: recorded under the networkx installed in this container (3.x, version string patched to the 2.2 the reference
demands, oracle/ref_loader.py); 2.2's sub-graph views order hub nodes' neighbours differently, which can only change
the ORDER of the induced pairs, not the sampled drugs or the pair set.

Usage:  python oracle/make_golden_samplers.py [--out tests/golden]"""
import argparse
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    args = ap.parse_args()
    ref_loader.load_reference()
    train_data, _, _, FLAGS = ref_loader.load_drugbank_fold(1)
    from sampler import NeighborSampler, EverythingSampler, RandomSampler
    from utils.util import set_seed
    G = train_data.dataset.interaction_combo_nxgraph
    assert all(list(G.neighbors(u)) == sorted(G.neighbors(u)) for u in G.nodes)      # adjacency order = ascending
    set_seed(8)
    s = NeighborSampler(train_data, 5, 64)
    out = {}
    for i in range(6):
        bg, sg, sub = s.sample_next_training_batch()
        out['batch_gids/%d' % i] = np.asarray(bg, np.int64)
        out['sampled_gids/%d' % i] = np.asarray(sg, np.int64)
        out['sub_nodes/%d' % i] = np.asarray(list(sub.nodes), np.int64)
    out['visited_counter'] = s.nodes_visited_counter.copy()
    set_seed(8)
    e = EverythingSampler(train_data)
    bg, sg, _ = e.sample_next_training_batch()
    out['everything/batch_gids_head'] = np.asarray(bg[:256], np.int64)
    out['everything/n'] = np.int64(len(bg))
    out['everything/sampled_n'] = np.int64(len(sg))
    set_seed(8)
    r = RandomSampler(train_data, 64, sample_induced=True)          # src/sampler.py:133-142
    for i in range(4):
        bg, sg, _ = r.sample_next_training_batch()
        out['induced/batch_gids/%d' % i] = np.asarray(bg, np.int64)
        out['induced/sampled_gids/%d' % i] = np.asarray(sg, np.int64)
    out['induced/visited_counter'] = r.nodes_visited_counter.copy()
    np.savez_compressed(os.path.join(args.out, 'bignn_samplers.npz'), **out)
    print('wrote', os.path.join(args.out, 'bignn_samplers.npz'))


if __name__ == '__main__':
    main()
