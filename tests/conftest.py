import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without CUDA skips the gpu-marked tests instead of failing at the first one."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def drugbank():
    from oracle.bignn_oracle import PackedDataset
    return PackedDataset.load(os.path.join(GOLDEN, 'drugbank_packed.npz'))


@pytest.fixture(scope='session')
def step_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, 'bignn_gin_gcn_step.npz'))


@pytest.fixture(scope='session')
def gin_gcn_specs():
    from oracle.bignn_oracle import parse_specs
    with open(os.path.join(GOLDEN, 'bignn_gin_gcn_layers.txt')) as f:
        return parse_specs(f.read().splitlines())
