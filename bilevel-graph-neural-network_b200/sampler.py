"""RandomSampler (src/sampler.py:110-131): shuffled mini-batches of positive train pairs
through the same torch DataLoader mechanism (same global-RNG consumption), so batches are
bit-identical to the reference's for an identical RNG state."""
import numpy as np
from torch.utils.data import DataLoader


class RandomSampler(object):
    def __init__(self, data, batch_size, sample_induced=False):
        if sample_induced:
            raise NotImplementedError('sample_induced is off by default (src/config.py) and not on the path')
        self.batch_size = batch_size
        self.data_loader = DataLoader(data, batch_size=batch_size, shuffle=True)
        self.data_iterable = iter(self.data_loader)

    def sample_next_training_batch(self):
        try:
            sampled_pairs = next(self.data_iterable)
        except StopIteration:
            self.data_iterable = iter(self.data_loader)
            sampled_pairs = next(self.data_iterable)
        batch_gids = sampled_pairs.cpu().detach().numpy()
        return batch_gids, np.unique(batch_gids), None
