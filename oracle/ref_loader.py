"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference (/root/reference) in this container on top of
the vendored third-party shim (oracle/shim), so that golden vectors can be
generated from the reference's own code.  This module can
only run where /root/reference exists (the build container); nothing on the GPU
box may import it.

What is patched, and why (none of it touches reference sources on disk):
  * sys.path gets oracle/shim first (torch_geometric / torch_scatter /
    torch_sparse / klepto / pytz / tensorboardX are not installable offline).
  * networkx.__version__ = '2.2' and nx.from_scipy_sparse_matrix (reference
    utils/util.py:20-28 hard-requires 2.2; utils/data/dataset.py:78,112).
  * `config` is executed from the reference's src/config.py text with the
    module-level selector lines replaced in memory (the layer specs
    are generated at import time from module variables, not argv).
  * utils.util.get_save_path -> a scratch dir (the tree is read-only);
    utils.util.load of the missing graph_data.klepto blob -> zeros(64) ECFP
    features (only used by init_embds='graph_feats').
"""
import os
import sys
import types

REF = os.environ.get('BIGNN_REFERENCE', '/root/reference')
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')
SCRATCH = os.environ.get('BIGNN_REF_SCRATCH', '/tmp/bignn_ref_scratch')


def _fake_graph_data(klepto_path):
    """graph_data.klepto stand-in: {fname: {'drug_feat': {'64': zeros(64)}}} for
    every .gexf file of the dataset (the blob is listed in .MISSING_LARGE_BLOBS;
    it only supplies ECFP `drug_feat`, used by init_embds='graph_feats')."""
    import glob
    import numpy as np
    root = os.path.dirname(os.path.dirname(klepto_path))
    files = glob.glob(os.path.join(root, 'ddi_data', 'drugs_snap', '*.gexf')) + \
        glob.glob(os.path.join(root, 'ddi_data', 'drugs_small', '*.gexf')) + \
        glob.glob(os.path.join(root, '*.gexf'))
    return {os.path.basename(f).split('.')[0]: {'drug_feat': {'64': np.zeros(64)}}
            for f in files}


def load_reference(model='lower_level_gnn_higher_level', lower='gin', higher='gcn',
                   dataset='drugbank', gpu=-1, extra_replacements=()):
    """Returns the reference's `config` module (with FLAGS) after making the
    reference importable.  Must be called once per process, before anything
    from the reference is imported."""
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree not present at {}'.format(REF))
    import networkx as nx
    nx.__version__ = '2.2'
    if not hasattr(nx, 'from_scipy_sparse_matrix'):
        nx.from_scipy_sparse_matrix = nx.from_scipy_sparse_array
    for p in (os.path.join(REF, 'src'), REF, SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.argv = ['main.py']
    os.makedirs(SCRATCH, exist_ok=True)

    import utils.util as uu
    uu.get_save_path = lambda: os.path.join(SCRATCH, 'save')
    _orig_load = uu.load

    def _load(filepath, print_msg=True):
        if filepath.endswith(os.path.join('klepto', 'graph_data.klepto')):
            return _fake_graph_data(filepath)
        return _orig_load(filepath, print_msg)
    uu.load = _load
    uu.get_root_path_for_logs = lambda: SCRATCH

    src = open(os.path.join(REF, 'src', 'config.py')).read()
    reps = [
        ("gpu = 1  # -1 if use cpu", "gpu = {}".format(gpu)),
        ("\ndataset = 'drugbank'\n", "\ndataset = '{}'\n".format(dataset)),
        ("\nmodel = 'higher_level_gnn'                  # DECAGON\n",
         "\nmodel = '{}'\n".format(model)),
        ("lower_level_gnn_type = 'gat'", "lower_level_gnn_type = '{}'".format(lower)),
        ("higher_level_gnn_type = 'gat'", "higher_level_gnn_type = '{}'".format(higher)),
    ] + list(extra_replacements)
    for a, b in reps:
        if a not in src:
            raise RuntimeError('config.py selector line not found: {!r}'.format(a))
        src = src.replace(a, b, 1)
    mod = types.ModuleType('config')
    mod.__file__ = os.path.join(REF, 'src', 'config.py')
    sys.modules['config'] = mod
    exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    return mod


def load_drugbank_fold(fold=1):
    """Runs the reference's own data path (src/main.py:24-46) for one fold and
    returns (train_data, val_pairs, test_pairs, FLAGS)."""
    from copy import deepcopy
    from config import FLAGS
    from load_data import load_pair_tvt_splits, load_pairs_to_dataset, load_dataset
    from utils.data.interaction_edge_feat import encode_edge_features
    from utils.data.representation_node_feat import encode_node_features
    from utils.util import set_seed

    tvt = load_pair_tvt_splits()
    orig = load_dataset(FLAGS.dataset, 'all', FLAGS.node_feats, FLAGS.edge_feats)
    orig, num_node_feat = encode_node_features(dataset=orig)
    nief = encode_edge_features(orig.interaction_combo_nxgraph, FLAGS.hyper_eatts)
    i = fold - 1
    set_seed(FLAGS.random_seed + 5)
    dataset = deepcopy(orig)
    train_data, val_data, test_data, val_pairs, test_pairs, _ = load_pairs_to_dataset(
        num_node_feat, nief, tvt['train'][i], tvt['val'][i], tvt['test'][i], dataset)
    return train_data, val_pairs, test_pairs, FLAGS


def prepare_drugcombo_subset():
    """DrugCombo as far as the reference tree still holds it.  Missing (listed in .MISSING_LARGE_BLOBS): the raw
    interaction table `ddi_data/Syner&Antag_voting.csv`, `klepto/graph_data.klepto` and (unreadable without klepto)
    `klepto/drug_name_to_cid`.  Present: the 3 376 molecule graphs `CIDs%08d.gexf` and the two typed interaction
    graphs `ddi_graphs/{synergy,antagonism}_ddi.gexf` whose node attribute `gid` is the CID.  This
    writes, into the scratch directory, a data tree the reference's own loader (utils/data/load_raw_data.py:34-58)
    accepts: symlinks to the molecule graphs and an interaction table with one row per typed edge whose two drugs
    both have a molecule graph (1 621 drugs, 11 410 rows; drugs without a graph are dropped by the reference itself,
    load_raw_data.py:144-156).  Returns (data_path, drug-name -> CID map)."""
    import networkx as nx
    src = os.path.join(REF, 'data', 'DrugCombo')
    data = os.path.join(SCRATCH, 'data')
    dc = os.path.join(data, 'DrugCombo')
    os.makedirs(os.path.join(dc, 'ddi_data'), exist_ok=True)
    os.makedirs(os.path.join(dc, 'klepto'), exist_ok=True)
    for f in sorted(os.listdir(src)):
        if f.endswith('.gexf') and not os.path.lexists(os.path.join(dc, f)):
            os.symlink(os.path.join(src, f), os.path.join(dc, f))
    rows, names = [], {}
    for label in ('synergy', 'antagonism'):
        g = nx.read_gexf(os.path.join(src, 'ddi_graphs', label + '_ddi.gexf'))
        gid = {u: int(d['gid']) for u, d in g.nodes(data=True)}
        for u, v in g.edges():
            fa, fb = 'CIDs%08d' % gid[u], 'CIDs%08d' % gid[v]
            if os.path.exists(os.path.join(src, fa + '.gexf')) and os.path.exists(os.path.join(src, fb + '.gexf')):
                rows.append((fa.lower(), fb.lower(), label))
                names[fa.lower()], names[fb.lower()] = fa, fb
    with open(os.path.join(dc, 'ddi_data', 'Syner&Antag_voting.csv'), 'w') as f:
        f.write('idx,drug1,drug2,label\n')                       # parse_edges_drugcombo reads columns 1, 2 and -1
        for i, (a, b, l) in enumerate(rows):
            f.write('%d,%s,%s,%s\n' % (i, a, b, l))
    return data, names


def patch_for_drugcombo():
    """after load_reference(dataset='drugcombo', ...): point the reference at the scratch data tree."""
    import utils.util as uu
    data, names = prepare_drugcombo_subset()
    uu.get_data_path = lambda: data
    prev = uu.load

    def _load(filepath, print_msg=True):
        if filepath.endswith(os.path.join('klepto', 'drug_name_to_cid')):
            return names
        return prev(filepath, print_msg)
    uu.load = _load
