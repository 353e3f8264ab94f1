"""Evaluation path on the GPU against the reference's own `evaluate` (src/train.py:185-220) output, recorded right
after the golden train step for 512 validation pairs (tests/golden/bignn_gin_gcn_eval.npz, oracle/make_golden.py
--eval_only 512): model.eval() statistics, the step's init_x, ONE upper pass and ONE scorer launch over all pairs
(`BiGNNEngine.score_pairs`) instead of an upper pass and 128 `.item()` syncs per 64-pair batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bignn_b200 as B
from bignn_b200.engine import BiGNNEngine

DEV = 'cuda:0'


def test_score_pairs_matches_reference_evaluate(golden_dir, step_golden):
    B._lib.load()
    z, e = step_golden, np.load(os.path.join(golden_dir, 'bignn_gin_gcn_eval.npz'))
    B.set_flags(B.make_flags(device=DEV))
    data = B.BiGNNData.from_npz(os.path.join(golden_dir, 'drugbank_packed.npz'), device=DEV)
    model = B.Model(data).to(DEV)
    sd = {k[4:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd1/')}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    model.train()
    eng = BiGNNEngine(data, model, use_cuda_graph=False)
    data.interaction_combo_nxgraph.init_x = torch.from_numpy(z['init_x']).to(DEV)
    got = eng.score_pairs(e['gids'], recompute_init_x=False)
    assert got.shape == (512, 1) and model.training
    err = float((got.view(-1).cpu().double() - torch.from_numpy(e['preds']).view(-1).double()).abs().max()
                / np.abs(e['preds']).max())
    assert err < 1e-5, err
    # the same predictions give the reference's mean validation loss (BCE per 64-pair batch, averaged)
    p = got.view(-1).cpu().double().clamp(1e-12, 1 - 1e-12)
    y = torch.from_numpy(e['y_true']).double()
    losses, off = [], 0
    for n in e['batch_sizes'].tolist():
        pp, yy = p[off:off + n], y[off:off + n]
        losses.append(float(-(yy * pp.log() + (1 - yy) * (1 - pp).log()).mean()))
        off += n
    assert abs(np.mean(losses) - float(e['mean_loss'])) < 1e-5
