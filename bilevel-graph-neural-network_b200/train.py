"""Step driver with the reference's sequencing (src/train.py:48-108,177-182): the all-drug
lower pass in 128-graph chunks, the sampled pair batch, the upper pass, backward, Adam.
This is the layer-by-layer drop-in path; `engine.BiGNNEngine` runs the same arithmetic as
one batched, CUDA-graph-captured step."""
import numpy as np
import torch

from .batch import BatchData
from .config import get_flags
from .sampler import RandomSampler


def all_drug_chunks(gids, batch_size):
    """Pair/chunk schedule of src/train.py:52-71 -> list of [P_c,2] gid arrays."""
    gids = list(gids)
    pairs = [(gids[i], gids[i + 1]) for i in range(0, len(gids) - 2, 2)]
    pairs.append((gids[-2], gids[-1]))
    pairs = np.asarray(pairs, np.int64)
    bs = int(len(pairs) / 2) if batch_size * 2 >= len(pairs) else batch_size
    out, i = [], 0
    for i in range(0, pairs.shape[0] - bs, bs):
        out.append(pairs[i:i + bs])
    out.append(pairs[i + bs:])
    return out


def _get_initial_embd(train_data, model):
    flags = get_flags()
    model.train_data = train_data
    model.use_layers = 'lower_layers'
    ds = train_data.dataset
    ig = ds.interaction_combo_nxgraph
    out = []
    for pairs in all_drug_chunks(list(ds.gs_map.keys()), flags.batch_size):
        bd = BatchData(pairs, ds, is_train=False, ignore_pairs=True)
        if ig.init_x is None:
            width = sum(l.out_dim for l in model.init_layers) if model.lower_layers[-1].concat_multi_scale \
                else model.init_layers[-1].out_dim
            ig.init_x = torch.zeros((ds.N, width), dtype=torch.float32, device=ds.device)
        out.append(model(bd))
    return torch.cat(out, dim=0)


def model_forward(model, data, sampler=None, is_train=True):
    flags = get_flags()
    if sampler is None:
        sampler = RandomSampler(data, flags.batch_size, flags.sample_induced)
    both = flags.lower_level_layers and flags.higher_level_layers
    if both and 'model_init' in flags.init_embds and is_train:
        _get_initial_embd(data, model)
    batch_gids, sampled_gids, subgraph = sampler.sample_next_training_batch()
    bd = BatchData(batch_gids, data.dataset, is_train=is_train, sampled_gids=sampled_gids,
                   enforce_negative_sampling=flags.enforce_negative_sampling,
                   unique_graphs=flags.batch_unique_graphs, subgraph=subgraph,
                   merge_graphs=not both or flags.pair_interaction)
    if both:
        model.use_layers = 'higher_layers'
    return bd


def _train_iter(batch_data, model, optimizer):
    loss = model(batch_data)
    loss.backward()
    optimizer.step()
    return loss.item()


def train_step(model, data, sampler, optimizer):
    """One iteration of the reference's hot loop (src/train.py:137-141)."""
    model.train()
    model.zero_grad()
    bd = model_forward(model, data, sampler=sampler)
    loss = _train_iter(bd, model, optimizer)
    bd.restore_interaction_nxgraph()
    return loss, bd
