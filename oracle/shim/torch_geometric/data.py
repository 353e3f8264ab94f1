"""`torch_geometric.data.Data` (1.1.2): attribute bag."""
import re

import torch

from .utils import is_undirected as _is_undirected


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def keys(self):
        return [key for key in self.__dict__.keys() if self[key] is not None]

    def __len__(self):
        return len(self.keys)

    def __contains__(self, key):
        return key in self.keys

    def __iter__(self):
        for key in sorted(self.keys):
            yield key, self[key]

    def __call__(self, *keys):
        for key in sorted(self.keys) if not keys else keys:
            if self[key] is not None:
                yield key, self[key]

    def __cat_dim__(self, key, value):
        return -1 if bool(re.search('(index|face)', key)) else 0

    def __cumsum__(self, key, value):
        return bool(re.search('(index|face)', key))

    @property
    def num_nodes(self):
        for key, item in self('x', 'pos'):
            return item.size(self.__cat_dim__(key, item))
        if self.edge_index is not None:
            return int(self.edge_index.max().item()) + 1
        return None

    @property
    def num_edges(self):
        for key, item in self('edge_index', 'edge_attr'):
            return item.size(self.__cat_dim__(key, item))
        return None

    @property
    def num_features(self):
        return 1 if self.x.dim() == 1 else self.x.size(1)

    def is_undirected(self):
        return _is_undirected(self.edge_index, self.num_nodes)

    def apply(self, func, *keys):
        for key, item in self(*keys):
            if torch.is_tensor(item):
                self[key] = func(item)
        return self

    def contiguous(self, *keys):
        return self.apply(lambda x: x.contiguous(), *keys)

    def to(self, device, *keys):
        return self.apply(lambda x: x.to(device), *keys)
