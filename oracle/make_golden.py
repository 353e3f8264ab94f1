"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* from the reference itself.

Runs the UNMODIFIED reference (/root/reference, via oracle/ref_loader.py + the
vendored third-party shim) on CPU in this container and records, for DrugBank
fold 1 with set_seed(3+5) (reference src/main.py:42):

  drugbank_packed.npz   the dataset as flat arrays (what utils/data/dataset.py
                        holds as networkx objects): molecule graphs, one-hot
                        features, train interaction graph, pair labels.
  bignn_gin_gcn_step.npz  one full Bi-GNN (GIN lower + GCN upper) train step:
                        initial state_dict, per-chunk merged-batch indexing,
                        activations, init_x, loss, every parameter gradient,
                        state_dict after the Adam step.
  sampler_seq.npz       24 consecutive steps of positive batches, negative
                        samples (order included) and labels.
(Known-answer micro cases -- P3 path, star, two components, isolated node,
duplicate / self-loop coalesce input -- are written out by hand in
tests/test_known_answers_cpu.py and tests/test_gpu_kernels.py.)

Usage:  python oracle/make_golden.py [--out tests/golden] [--config gin_gcn|gat_gat|...]
The reference cannot travel to the GPU box, hence the committed fixtures.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def pack_dataset(train_data, val_pairs, test_pairs):
    """Flat-array view of the reference's BiGNNDataset (utils/data/dataset.py:47)."""
    import torch
    from model.layers_util import convert_nx_to_pyg_graph
    ds = train_data.dataset
    gids = np.asarray(list(ds.gs_map.keys()), dtype=np.int64)
    assert list(ds.gs_map.values()) == list(range(len(gids)))
    atom_ptr = [0]
    xs, rows, cols, nbr_ptr = [], [], [], [0]
    for g in ds.gs:
        nxg = g.get_nxgraph()
        data = convert_nx_to_pyg_graph(nxg)       # reference model/layers_util.py:100
        n = data.x.shape[0]
        ei = data.edge_index.numpy()
        # sorted lexicographically by (row, col) -> CSR directly
        assert np.all(np.diff(ei[0] * n + ei[1]) > 0)
        cnt = np.bincount(ei[0], minlength=n)
        nbr_ptr.extend((nbr_ptr[-1] + np.cumsum(cnt)).tolist())
        cols.append(ei[1].astype(np.int32))
        xs.append(data.x.numpy())
        atom_ptr.append(atom_ptr[-1] + n)
    x = np.concatenate(xs, 0)
    assert np.all((x == 0) | (x == 1))
    hi = convert_nx_to_pyg_graph(ds.interaction_combo_nxgraph) \
        if not isinstance(ds.interaction_combo_nxgraph.init_x, list) else None
    # interaction graph edge list via the reference's own create_edge_index
    from model.layers_util import create_edge_index
    ddi, _ = create_edge_index(ds.interaction_combo_nxgraph)
    ddi = ddi.numpy()
    pair_keys = np.asarray(list(ds.pairs.keys()), dtype=np.int64)
    pair_labels = np.asarray([p.true_label for p in ds.pairs.values()], dtype=np.int8)
    out = dict(
        gids=gids,
        atom_ptr=np.asarray(atom_ptr, np.int32),
        x_u8=x.astype(np.uint8),
        nbr_ptr=np.asarray(nbr_ptr, np.int32),
        nbr_idx=np.concatenate(cols).astype(np.int32),
        ddi_row=ddi[0].astype(np.int32), ddi_col=ddi[1].astype(np.int32),
        train_pairs=train_data.data_items.numpy().astype(np.int64),
        pair_keys=pair_keys, pair_labels=pair_labels,
        val_pairs=val_pairs.numpy().astype(np.int64),
        test_pairs=test_pairs.numpy().astype(np.int64),
        num_labels=np.int64(ds.num_labels),
    )
    # per-edge-type interaction graphs (DrugCombo; utils/data/dataset.py:105-115), in the order
    # LoadInteractionGraph hands them to NodeModelAggrByEdge.GNNS[i] (model/layers_load_interaction_graph.py:16-19):
    # the index prefix keeps that order under sorting
    if getattr(ds, 'interaction_nxgraphs', None) and len(ds.interaction_nxgraphs) > 1:
        for i, (k, g) in enumerate(ds.interaction_nxgraphs.items()):
            ei = create_edge_index(g)[0].numpy()
            out['etype_row/%d_%s' % (i, k)] = ei[0].astype(np.int32)
            out['etype_col/%d_%s' % (i, k)] = ei[1].astype(np.int32)
    return out


def sd_to_np(sd):
    # canonical names only (the reference registers the same modules under five
    # ModuleList aliases, model/model.py:16-27)
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items() if k.startswith('layers.')}


def run_step_golden(train_data, FLAGS, out_path, n_seq=24, light=False, eval_pairs=None, eval_out=None):
    """One recorded train step (reference src/train.py:136-141) + sampler sequence."""
    import torch
    import train as T
    from batch import BatchData
    from model.model import Model
    from sampler import RandomSampler
    from utils.util import set_seed
    import model.layers_aggregation as LA

    ds = train_data.dataset
    set_seed(FLAGS.random_seed + 5)
    model = Model(train_data)
    sd0 = sd_to_np(model.state_dict())

    rec = {}
    # --- record the per-chunk merged batches of the all-drug pass -------------
    chunks = []
    orig_init = BatchData.__init__

    def rec_init(self, *a, **k):
        orig_init(self, *a, **k)
        if self.ignore_pairs:
            chunks.append(self)
    BatchData.__init__ = rec_init

    acts_per_chunk = []
    orig_fwd = Model.forward

    def rec_fwd(self, batch_data):
        out = orig_fwd(self, batch_data)
        if getattr(batch_data, 'ignore_pairs', False):
            acts_per_chunk.append([a.detach().numpy().copy() for a in self.acts])
        else:
            rec['upper_acts'] = [a.detach().numpy().copy() for a in self.acts]
        return out
    Model.forward = rec_fwd

    # reference train(): initial fill (src/train.py:22-28)
    T._get_initial_embd(train_data, model)
    ds.init_interaction_graph_embds(device=FLAGS.device)
    n_init_chunks = len(chunks)
    init_bn = sd_to_np(model.state_dict())           # BN running stats moved by the initial fill
    opt = torch.optim.Adam(model.parameters(), lr=FLAGS.lr)
    sampler = RandomSampler(train_data, FLAGS.batch_size, FLAGS.sample_induced)

    del chunks[:]
    del acts_per_chunk[:]
    rng_state_np = np.random.get_state()
    model.train()
    model.zero_grad()
    bd = T.model_forward(model, train_data, sampler=sampler)
    init_x = ds.interaction_combo_nxgraph.init_x.detach().numpy().copy()
    loss = model(bd)
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()
             if k.startswith('layers.') and p.grad is not None}
    preds = bd.pair_list and np.asarray([float(np.asarray(p.link_pred).reshape(-1)[0])
                                         for p in bd.pair_list], np.float32)
    opt.step()
    bd.restore_interaction_nxgraph()
    sd1 = sd_to_np(model.state_dict())

    rec.update(
        n_chunks=np.int64(len(chunks)), loss=np.float32(loss.item()),
        init_x=init_x, pair_preds=preds,
        batch_gids=np.asarray(bd.batch_gids, np.int64),
        positive_gids=np.asarray(bd.positive_pair_gids, np.int64),
        negative_gids=np.asarray(bd.negative_pair_gids, np.int64),
        y_true=np.asarray([p.true_label for p in bd.pair_list], np.int64),
        sampled_gids=np.asarray(bd.sampled_gids, np.int64),
    )
    for k, v in sd0.items():
        rec['sd0/' + k] = v
    for k, v in init_bn.items():
        if 'running' in k or 'num_batches' in k:
            rec['sd_init/' + k] = v
    for k, v in sd1.items():
        if light and not ('running' in k or 'num_batches' in k or k.startswith('layers.%d.' % (FLAGS.layer_num - 2))):
            continue          # light fixture: post-step BatchNorm buffers and the scorer only
        rec['sd1/' + k] = v
    for k, v in grads.items():
        rec['grad/' + k] = v
    # per-chunk indexing (bit-exact targets) -- all chunks; activations: first & last chunk
    for c, b in enumerate(chunks):
        if light:             # indexing is pinned by the DrugBank fixture; keep the chunk schedule only
            rec['chunk%d/gids' % c] = np.asarray(list(b.merge_data['gids_to_batch_ind'].keys()), np.int64)
            continue
        md = b.merge_data
        m = md['merge']
        rec['chunk%d/gids' % c] = np.asarray(list(md['gids_to_batch_ind'].keys()), np.int64)
        rec['chunk%d/batch_gids' % c] = np.asarray(b.batch_gids, np.int64)
        rec['chunk%d/ind_list' % c] = np.asarray(md['ind_list'], np.int64)
        rec['chunk%d/edge_ind_list' % c] = np.asarray(md['edge_ind_list'], np.int64)
        rec['chunk%d/graph_sizes' % c] = np.asarray(md['graph_sizes'], np.int64)
        if c in (0, len(chunks) - 1):
            rec['chunk%d/edge_index' % c] = m.edge_index.numpy().astype(np.int32)
            rec['chunk%d/batch' % c] = m.batch.numpy().astype(np.int32)
            rec['chunk%d/x_u8' % c] = m.x.numpy().astype(np.uint8)
        if c == len(chunks) - 1:
            for li, a in enumerate(acts_per_chunk[c][1:]):
                rec['chunk%d/act%d' % (c, li + 1)] = a
        if c == 0:
            rec['chunk0/pooled'] = acts_per_chunk[0][-1]
    rec['upper_acts_n'] = list(range(len(rec['upper_acts'])))
    for li, a in enumerate(rec.pop('upper_acts')):
        if li == 0:
            continue          # acts[0] is the pair batch's merged x (unused upstairs)
        if a.ndim == 0:
            continue
        if light and li not in (2, len(rec['upper_acts_n']) - 3, len(rec['upper_acts_n']) - 2):
            continue          # light fixture: first and last upper-level activations + scorer output (act1 = init_x)
        rec['upper/act%d' % li] = a
    rec.pop('upper_acts_n', None)
    if eval_out is not None:
        # --- evaluation path (reference src/train.py:185-220) right after the recorded step: model.eval(), the
        # init_x of that step, one upper pass + scorer per 64-pair batch.  Written to its own file; the step and
        # sampler fixtures are NOT rewritten in this mode (evaluate() consumes torch RNG for its shuffle).
        eval_batches = []

        def rec_eval_init(self, *a, **k):
            orig_init(self, *a, **k)
            if not self.ignore_pairs and not self.is_train:
                eval_batches.append(self)
        BatchData.__init__ = rec_eval_init
        pl, eloss = T.evaluate(model, train_data, eval_pairs, None)
        BatchData.__init__ = orig_init
        Model.forward = orig_fwd
        gids = np.concatenate([np.asarray(b.batch_gids, np.int64) for b in eval_batches])
        preds = np.asarray([np.asarray(p.link_pred, np.float32).reshape(-1) for b in eval_batches for p in b.pair_list])
        ys = np.asarray([p.true_label for b in eval_batches for p in b.pair_list], np.int64)
        ev = dict(gids=gids, preds=preds, y_true=ys, mean_loss=np.float32(eloss),
                  batch_sizes=np.asarray([len(b.pair_list) for b in eval_batches], np.int64), init_x=init_x)
        for k, v in sd1.items():
            ev['sd1/' + k] = v
        np.savez_compressed(eval_out, **ev)
        print('wrote', eval_out, 'pairs', gids.shape, 'mean loss', eloss)
        return None
    np.savez_compressed(out_path, **rec)
    print('wrote', out_path, 'loss', loss.item(), 'chunks', len(chunks), 'init chunks', n_init_chunks)

    # --- sampler sequence: continue for n_seq more steps ----------------------
    seq = dict(np_state_keys=np.asarray(rng_state_np[1], np.uint32),
               np_state_pos=np.int64(rng_state_np[2]))
    pos, neg, ys, losses = [], [], [], []
    if n_seq == 0:
        BatchData.__init__ = orig_init
        Model.forward = orig_fwd
        return seq
    for it in range(n_seq):
        model.train()
        model.zero_grad()
        bd = T.model_forward(model, train_data, sampler=sampler)
        l = T._train_iter(bd, model, opt)
        bd.restore_interaction_nxgraph()
        pos.append(np.asarray(bd.positive_pair_gids, np.int64))
        neg.append(np.asarray(bd.negative_pair_gids, np.int64))
        ys.append(np.asarray([p.true_label for p in bd.pair_list], np.int64))
        losses.append(l)
    seq.update(first_pos=rec['positive_gids'], first_neg=rec['negative_gids'], first_y=rec['y_true'],
               pos=np.stack(pos), neg=np.stack(neg), y=np.stack(ys),
               losses=np.asarray(losses, np.float32))
    BatchData.__init__ = orig_init
    Model.forward = orig_fwd
    return seq


def run_lower_only_golden(train_data, FLAGS, out_path):
    """One recorded train step of the lower-level-only model (model='lower_level_gnn', the LL-GNN baseline; SURVEY 3.5
    / BASELINE config 3): src/train.py:99-107 builds ONE BatchData of the pair batch's unique molecule graphs,
    Model.forward runs 5 x GIN -> multi-scale readout [G, 320] -> LinkPred over gids_to_batch_ind rows
    (MLP 640-80-10-1, model/layers_link_pred.py:52) -> BCE."""
    import torch
    import train as T
    from model.model import Model
    from sampler import RandomSampler
    from utils.util import set_seed
    set_seed(FLAGS.random_seed + 5)
    model = Model(train_data)
    sd0 = sd_to_np(model.state_dict())
    opt = torch.optim.Adam(model.parameters(), lr=FLAGS.lr)
    sampler = RandomSampler(train_data, FLAGS.batch_size, FLAGS.sample_induced)
    model.train()
    model.zero_grad()
    bd = T.model_forward(model, train_data, sampler=sampler)
    loss = model(bd)
    acts = [a.detach().numpy().copy() for a in model.acts]
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()
             if k.startswith('layers.') and p.grad is not None}
    opt.step()
    sd1 = sd_to_np(model.state_dict())
    md = bd.merge_data
    rec = dict(loss=np.float32(loss.item()),
               batch_gids=np.asarray(bd.batch_gids, np.int64),
               positive_gids=np.asarray(bd.positive_pair_gids, np.int64),
               negative_gids=np.asarray(bd.negative_pair_gids, np.int64),
               sampled_gids=np.asarray(bd.sampled_gids, np.int64),
               y_true=np.asarray([p.true_label for p in bd.pair_list], np.int64),
               merge_gids=np.asarray(list(md['gids_to_batch_ind'].keys()), np.int64),
               ind_list=np.asarray(md['ind_list'], np.int64),
               edge_index=md['merge'].edge_index.numpy().astype(np.int32),
               batch=md['merge'].batch.numpy().astype(np.int32))
    # acts: [x, 5 x NodeEmbedding, NodeAggregation [G, 320], LinkPred [P, 1], loss]
    for li in (1, 5, 6, 7):
        rec['act%d' % li] = acts[li]
    for k, v in sd0.items():
        rec['sd0/' + k] = v
    for k, v in sd1.items():
        if 'running' in k or 'num_batches' in k:
            rec['sd1/' + k] = v
    for k, v in grads.items():
        rec['grad/' + k] = v
    np.savez_compressed(out_path, **rec)
    print('wrote', out_path, 'loss', loss.item(), 'unique graphs', len(rec['merge_gids']), 'atoms', acts[1].shape[0])


def run_upper_only_golden(train_data, FLAGS, out_path):
    """One recorded train step of the upper-level-only model (model='higher_level_gnn', DECAGON -- the model the
    reference's config.py selects as shipped): fixed random drug features (init_embds='rand_init',
    utils/data/dataset.py:181-188) -> LoadInteractionLayer -> 3 x NodeEmbedding over the interaction graph ->
    LinkPred -> BCE."""
    import torch
    import train as T
    from model.model import Model
    from sampler import RandomSampler
    from utils.util import set_seed
    ds = train_data.dataset
    set_seed(FLAGS.random_seed + 5)
    model = Model(train_data)
    sd0 = sd_to_np(model.state_dict())
    opt = torch.optim.Adam(model.parameters(), lr=FLAGS.lr)
    sampler = RandomSampler(train_data, FLAGS.batch_size, FLAGS.sample_induced)
    ds.init_interaction_graph_embds(device=FLAGS.device)          # src/train.py:22-28 (no lower level: graph_feats)
    feats = ds.interaction_combo_nxgraph.init_x.detach().numpy().copy()
    model.train()
    model.zero_grad()
    bd = T.model_forward(model, train_data, sampler=sampler)
    loss = model(bd)
    acts = [a.detach().numpy().copy() if a is not None else None for a in model.acts]
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()
             if k.startswith('layers.') and p.grad is not None}
    opt.step()
    sd1 = sd_to_np(model.state_dict())
    rec = dict(loss=np.float32(loss.item()), graph_feats=feats,
               batch_gids=np.asarray(bd.batch_gids, np.int64),
               positive_gids=np.asarray(bd.positive_pair_gids, np.int64),
               sampled_gids=np.asarray(bd.sampled_gids, np.int64),
               y_true=np.asarray([p.true_label for p in bd.pair_list], np.int64))
    for li in range(1, len(acts)):
        if acts[li] is not None and acts[li].ndim > 0:
            rec['act%d' % li] = acts[li]
    for k, v in sd0.items():
        rec['sd0/' + k] = v
    for k, v in sd1.items():
        if 'running' in k or 'num_batches' in k:
            rec['sd1/' + k] = v
    for k, v in grads.items():
        rec['grad/' + k] = v
    np.savez_compressed(out_path, **rec)
    print('wrote', out_path, 'loss', loss.item(), 'acts', [None if a is None else a.shape for a in acts])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    ap.add_argument('--lower', default='gin')
    ap.add_argument('--higher', default='gcn')
    ap.add_argument('--model', default='lower_level_gnn_higher_level')
    ap.add_argument('--tag', default='bignn_gin_gcn')
    ap.add_argument('--skip_pack', action='store_true')
    ap.add_argument('--dataset', default='drugbank', choices=['drugbank', 'drugcombo'])
    ap.add_argument('--n_seq', type=int, default=24)
    ap.add_argument('--light', action='store_true', help='smaller step fixture (no per-chunk indexing arrays)')
    ap.add_argument('--eval_only', type=int, default=0,
                    help='record the evaluation path on the first N validation pairs after the step -> <tag>_eval.npz '
                         '(writes nothing else)')
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    ref_loader.load_reference(model=args.model, lower=args.lower, higher=args.higher, dataset=args.dataset)
    if args.dataset == 'drugcombo':
        ref_loader.patch_for_drugcombo()
    train_data, val_pairs, test_pairs, FLAGS = ref_loader.load_drugbank_fold(1)
    if 'lower_level' not in args.model:
        run_upper_only_golden(train_data, FLAGS, os.path.join(args.out, args.tag + '_step.npz'))
        with open(os.path.join(args.out, args.tag + '_layers.txt'), 'w') as f:
            f.write('\n'.join(getattr(FLAGS, 'layer_%d' % i) for i in range(1, FLAGS.layer_num + 1)) + '\n')
        return
    if 'higher_level' not in args.model:
        run_lower_only_golden(train_data, FLAGS, os.path.join(args.out, args.tag + '_step.npz'))
        with open(os.path.join(args.out, args.tag + '_layers.txt'), 'w') as f:
            f.write('\n'.join(getattr(FLAGS, 'layer_%d' % i) for i in range(1, FLAGS.layer_num + 1)) + '\n')
        return
    if args.eval_only:
        run_step_golden(train_data, FLAGS, None, eval_pairs=val_pairs[:args.eval_only],
                        eval_out=os.path.join(args.out, args.tag + '_eval.npz'))
        return
    if not args.skip_pack:
        packed = pack_dataset(train_data, val_pairs, test_pairs)
        np.savez_compressed(os.path.join(args.out, args.dataset + '_packed.npz'), **packed)
        print('packed', {k: getattr(v, 'shape', v) for k, v in packed.items()})
    seq = run_step_golden(train_data, FLAGS, os.path.join(args.out, args.tag + '_step.npz'), n_seq=args.n_seq, light=args.light)
    np.savez_compressed(os.path.join(args.out, args.tag + '_sampler_seq.npz'), **seq)
    layer_specs = [getattr(FLAGS, 'layer_%d' % i) for i in range(1, FLAGS.layer_num + 1)]
    with open(os.path.join(args.out, args.tag + '_layers.txt'), 'w') as f:
        f.write('\n'.join(layer_specs) + '\n')


if __name__ == '__main__':
    main()
