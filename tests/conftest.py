import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_attrs(g):
    """sampler attributes stored in an ags_*.npz fixture (NaN encodes None)."""
    out = {}
    for k in g.files:
        if k.startswith('attr_'):
            v = g[k]
            out[k[5:]] = None if (v.dtype.kind == 'f' and np.isnan(float(v))) else v.item()
    return out


@pytest.fixture(scope='session')
def golden():
    return load_golden
