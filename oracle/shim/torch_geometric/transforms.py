"""`torch_geometric.transforms.LocalDegreeProfile` -- imported by the reference
(utils/data/representation_node_feat.py:4) but off by default (node_fe_1='one_hot')."""


class LocalDegreeProfile(object):
    def __call__(self, data):
        raise NotImplementedError('LocalDegreeProfile is outside the oracle scope')
